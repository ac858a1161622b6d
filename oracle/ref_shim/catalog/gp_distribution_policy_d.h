/* stand-in for the genbki-generated catalog/gp_distribution_policy_d.h: nothing from it is needed by the block-format sources */
