/* linux port header stand-in */
#define _GNU_SOURCE 1
