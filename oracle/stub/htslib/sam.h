/* Opaque stand-ins for the htslib types named by the reference's include/bs_call.h.
 * Test infrastructure only: lets the reference's hot-path .c files compile where htslib is absent.
 * None of the hot-path files dereference these types. */
#ifndef BSGPU_STUB_HTS_SAM_H
#define BSGPU_STUB_HTS_SAM_H
#include <stdint.h>
typedef struct htsFile htsFile;
typedef struct bam_hdr_t bam_hdr_t;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;
typedef struct bam1_t bam1_t;
#endif
