/* minihts -- bgzf subset: buffered byte input over a plain file or a BGZF (RFC 1952 members with the BC extra field, SAM
 * specification section 4.1) file, and BGZF block output.  src/read_reference.c reads the FASTA through bgzf_useek /
 * bgzf_getc (uncompressed files only here). */
#ifndef MINIHTS_BGZF_H
#define MINIHTS_BGZF_H
#include <stdint.h>
#include <sys/types.h>
typedef struct BGZF BGZF;
ssize_t bgzf_read(BGZF *fp, void *data, size_t length);      /* bytes read (less than length only at the end of the file), < 0 on error */
int bgzf_getc(BGZF *fp);
int bgzf_useek(BGZF *fp, off_t uoffset, int where);
#endif
