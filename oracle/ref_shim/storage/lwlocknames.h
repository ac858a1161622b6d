/* stand-in for storage/lwlocknames.h, which the reference build generates from storage/lmgr/lwlocknames.txt
 * (generate-lwlocknames.pl): only the count of individually named locks is needed by storage/lwlock.h's enum; the last
 * entry of lwlocknames.txt is CommittedGxidArrayLock 69 */
#ifndef LWLOCKNAMES_H
#define LWLOCKNAMES_H
#define NUM_INDIVIDUAL_LWLOCKS 70
#endif
