/* minihts -- hts.h subset (see ../minihts.c). */
#ifndef MINIHTS_HTS_H
#define MINIHTS_HTS_H
#include <stdint.h>
#include "kstring.h"
#include "bgzf.h"
#include "hfile.h"
typedef int64_t hts_pos_t;
enum htsFormatCategory { unknown_category, sequence_data, variant_data, index_file, region_list };
enum htsExactFormat { unknown_format, binary_format, text_format, sam, bam, bai, cram, crai, vcf, bcf, csi, gzi, tbi, bed, htsget, empty_format, fasta_format, fastq_format, fai_format, fqi_format };
enum htsCompression { no_compression, gzip, bgzf, custom, bzip2_compression };
typedef struct htsFormat {
	enum htsFormatCategory category;
	enum htsExactFormat format;
	struct { short major, minor; } version;
	enum htsCompression compression;
	short compression_level;
	void *specific;
} htsFormat;
typedef struct htsFile {
	uint32_t is_bin:1, is_write:1, is_be:1, is_cram:1, is_bgzf:1, dummy:27;
	int64_t lineno;
	kstring_t line;
	char *fn, *fn_aux;
	union { BGZF *bgzf; void *any; } fp;
	void *state;
	htsFormat format;
} htsFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;
htsFile *hts_open(const char *fn, const char *mode);
htsFile *hts_hopen(hFILE *fp, const char *fn, const char *mode);
int hts_close(htsFile *fp);
int hts_set_threads(htsFile *fp, int n);
int hts_set_fai_filename(htsFile *fp, const char *fn_aux);
void hts_idx_destroy(hts_idx_t *idx);
void hts_itr_destroy(hts_itr_t *itr);
#endif
