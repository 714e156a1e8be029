/* oracle/stub/lapacke_utils.h -- locate.c only takes MIN/MAX from this LAPACKE header. */
#ifndef ORACLE_STUB_LAPACKE_UTILS_H
#define ORACLE_STUB_LAPACKE_UTILS_H
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif
