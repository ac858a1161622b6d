/* stand-in for the generated utils/fmgroids.h (the reference build generates it; the files compiled or type-checked against
 * these stand-ins use none of its constants) */
#ifndef STANDIN_FMGROIDS_H
#define STANDIN_FMGROIDS_H
#endif
