/* minihts -- hfile subset: a file descriptor handed to hts_hopen() (src/process.c:128-129). */
#ifndef MINIHTS_HFILE_H
#define MINIHTS_HFILE_H
typedef struct hFILE hFILE;
hFILE *hdopen(int fd, const char *mode);
#endif
