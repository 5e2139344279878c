/* Build stub for oracle/_ref only: the reference's SimplePileupViewer.h:3 includes
 * htslib/faidx.h for one opaque pointer type. htslib is absent from this image and
 * the BAM half of the viewer is not compiled (see oracle/Makefile). */
#ifndef VB2_ORACLE_STUB_FAIDX_H
#define VB2_ORACLE_STUB_FAIDX_H
typedef struct faidx_t faidx_t;
#endif
