/* stand-in for the genbki-generated catalog/pg_class_d.h: nothing from it is needed by the block-format sources */
