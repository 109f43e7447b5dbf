/*
 * METIS 5 interface as the reference expects it (32-bit idx_t, TACSCreator.cpp:1105-1124,
 * TACSAssembler.cpp:1620-1624), adapted in metis_shim.c onto the 64-bit-idx_t
 * libmetis_static.a that ships with the CUDA toolkit.  Test infrastructure.
 */
#ifndef TACSB200_ORACLE_METIS_H
#define TACSB200_ORACLE_METIS_H
#define METIS_NOPTIONS 40
#define METIS_OPTION_NUMBERING 17
#define METIS_OK 1
#ifdef __cplusplus
extern "C" {
#endif
int METIS_SetDefaultOptions(int *options);
int METIS_PartGraphRecursive(int *nvtxs, int *ncon, int *xadj, int *adjncy, int *vwgt, int *vsize,
                             int *adjwgt, int *nparts, float *tpwgts, float *ubvec, int *options,
                             int *edgecut, int *part);
int METIS_PartGraphKway(int *nvtxs, int *ncon, int *xadj, int *adjncy, int *vwgt, int *vsize,
                        int *adjwgt, int *nparts, float *tpwgts, float *ubvec, int *options,
                        int *edgecut, int *part);
int METIS_NodeND(int *nvtxs, int *xadj, int *adjncy, int *vwgt, int *options, int *perm, int *iperm);
#ifdef __cplusplus
}
#endif
#endif
