/*
 * mceik_b200_types.h -- data-model structs shared with host code.
 *
 * Field names, order and C types are those of the reference's include/mceik_struct.h
 * (lines 4-90) because the drop-in boundary passes these structs by pointer: layouts must
 * match byte for byte (checked by tests/test_abi.py against the offsets of an LP64 build).
 * When the reference header is already included, its definitions are used as they are.
 */
#ifndef MCEIK_B200_TYPES_H
#define MCEIK_B200_TYPES_H 1

#ifndef _mceik_struct_h__ /* guard macro of the reference's own header */
#define _mceik_struct_h__ 1

/* phase of a pick; also selects which of a station's two tables is used (P -> 0, S -> 1) */
enum pick_type_enum { P_PRIMARY_PICK = 1, S_PRIMARY_PICK = 2 };

/* Event catalogue in CSR form: event e owns picks obsPtr[e] .. obsPtr[e+1]-1. */
struct mceik_catalog_struct {
    double *xsrc, *ysrc, *zsrc; /* [nevents] event position (m); z measured up from the model base */
    double *tori;               /* [nevents] origin time (epoch s) */
    double *tobs;               /* [npicks]  observed pick times (epoch s) */
    double *test;               /* [npicks]  predicted pick times (epoch s) */
    double *varObs;             /* [npicks]  pick variance (s) */
    int *luseObs;               /* [npicks]  0 = ignore this pick */
    int *pickType;              /* [npicks]  pick_type_enum */
    int *statPtr;               /* [npicks]  1-based station of the pick */
    int *obsPtr;                /* [nevents+1] CSR offsets */
    int nevents;
};

struct mceik_stations_struct {
    char **netw, **stnm, **chan, **loc; /* [nstat] SEED-style identifiers */
    double *xrec, *yrec, *zrec;         /* [nstat] station position (m) */
    double *pcorr, *scorr;              /* [nstat] P / S static corrections (s) */
    int *lhasP, *lhasS;                 /* [nstat] 1 if the station carries that phase */
    int nstat;
    int lcartesian; /* 1 = coordinates are Cartesian metres */
};

struct catalog_struct {
    int nevents;
};

struct mcmc_parms_struct {
    char resdir[512]; /* output directory */
    int nburnIn;      /* burn-in proposals */
    int niter;        /* total forward problems */
    int keepK;        /* thinning interval */
};

struct eik_parms_struct {
    double tol; /* convergence tolerance (s) */
    int maxit;  /* sweep-iteration cap */
};

struct mceik_parms_struct {
    struct mcmc_parms_struct mcparms;
    struct eik_parms_struct eikparms;
    char projnm[128];
    char scratch_dir[512];
    double x0, y0, z0;       /* model origin (m) */
    double dx, dy, dz;       /* node spacing (m) */
    int ndivx, ndivy, ndivz; /* domain divisions of the reference's MPI layout */
    int nrefx, nrefy, nrefz; /* inversion-grid -> eikonal-grid refinement factors */
};

#endif /* _mceik_struct_h__ */
#endif /* MCEIK_B200_TYPES_H */
