/* TEST INFRASTRUCTURE ONLY (oracle).  Minimal stand-in for <htslib/sam.h>.
 *
 * The reference's bwa/bwa.h:7-9 and src/bwa_wrapper.h:24-26,40-43 name a few
 * htslib types even when USE_HTSLIB is undefined; htslib itself is not in this
 * image.  Only the type names are needed: the text-SAM path
 * (src/bwa_wrapper.cpp:579-590 with USE_HTSLIB undefined) never touches them.
 */
#ifndef BWAFLOW_ORACLE_STUB_SAM_H
#define BWAFLOW_ORACLE_STUB_SAM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { int32_t tid, pos; uint16_t flag; } bam1_core_t;
typedef struct { bam1_core_t core; int l_data, m_data; uint8_t *data; } bam1_t;
typedef struct { int32_t n_targets; uint32_t l_text; char *text; } bam_hdr_t;
typedef struct samFile_s samFile;
bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
#ifdef __cplusplus
}
#endif
#endif
