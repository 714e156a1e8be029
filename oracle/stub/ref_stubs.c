/* oracle/stub/ref_stubs.c -- the reference declares weightedMedian__double (locate.c:73)
 * but defines it nowhere; only its L1 search and demo main use it.  A stub lets the
 * unmodified locate.c link as a shared object for the L2 path. */
#include <stdbool.h>
double weightedMedian__double(const int n, const double *x, const double *w, int *perm,
                              bool *lsort, int *ierr)
{
    (void)n; (void)x; (void)w; (void)perm; (void)lsort;
    if (ierr) *ierr = 1;
    return 0.0;
}
