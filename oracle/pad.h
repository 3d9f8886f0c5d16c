/* TEST INFRASTRUCTURE ONLY (oracle).  Force-included in front of the unmodified
 * reference sources when they are compiled into oracle/_ref/.
 * The reference's strrev() (src/alignment.h:172-184) writes s[l] one byte past a
 * calloc(l) block; padding every calloc keeps the reference alive on all inputs
 * without touching its source. */
#include <stdlib.h>
#define calloc(n, s) (calloc)((n) + 16, (s))
