/* Build stub for oracle/_ref only: samtools/sam_opts.h:31-38 embeds two htsFormat
 * values in sam_global_args. Only the struct's existence matters for the text path. */
#ifndef VB2_ORACLE_STUB_HTS_H
#define VB2_ORACLE_STUB_HTS_H
typedef struct htsFormat { int stub; } htsFormat;
#endif
