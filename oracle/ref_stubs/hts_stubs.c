/* TEST SCAFFOLDING for oracle/_ref only.
 * The oracle hand-encodes bam1_t records (ref_harness.cpp) so htslib is not built; these are our own
 * minimal stand-ins for the htslib entry points the reference objects link against.  Records built by
 * the harness carry no aux tags, so bam_aux_get() always answers "absent"; BAM/SAM file I/O aborts. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../htslib/htslib/hts.h"
#include "../htslib/htslib/sam.h"
#include "pod5_format/pod5_format_export.h"

#define DNB_UNREACHABLE(name) do { fprintf(stderr, "oracle/_ref: htslib stub %s called\n", name); abort(); } while (0)

bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) {
    if (!b) return;
    free(b->data);
    free(b);
}
bam1_t *bam_dup1(const bam1_t *src) {
    bam1_t *b = bam_init1();
    *b = *src;
    b->data = (uint8_t *)malloc(src->m_data ? src->m_data : 1);
    memcpy(b->data, src->data, (size_t)src->l_data);
    return b;
}
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) { (void)b; (void)tag; return NULL; }
int64_t bam_aux2i(const uint8_t *s) { (void)s; DNB_UNREACHABLE("bam_aux2i"); }
char *bam_aux2Z(const uint8_t *s) { (void)s; DNB_UNREACHABLE("bam_aux2Z"); }
uint32_t bam_auxB_len(const uint8_t *s) { (void)s; DNB_UNREACHABLE("bam_auxB_len"); }
int64_t bam_auxB2i(const uint8_t *s, uint32_t idx) { (void)s; (void)idx; DNB_UNREACHABLE("bam_auxB2i"); }
int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data) {
    (void)b; (void)tag; (void)type; (void)len; (void)data; DNB_UNREACHABLE("bam_aux_append");
}
int bam_aux_del(bam1_t *b, uint8_t *s) { (void)b; (void)s; DNB_UNREACHABLE("bam_aux_del"); }
int bam_aux_update_array(bam1_t *b, const char tag[2], uint8_t type, uint32_t items, void *data) {
    (void)b; (void)tag; (void)type; (void)items; (void)data; DNB_UNREACHABLE("bam_aux_update_array");
}
void bam_hdr_destroy(bam_hdr_t *h) { (void)h; }
bam_hdr_t *sam_hdr_read(samFile *fp) { (void)fp; DNB_UNREACHABLE("sam_hdr_read"); }
int sam_hdr_write(samFile *fp, const bam_hdr_t *h) { (void)fp; (void)h; DNB_UNREACHABLE("sam_hdr_write"); }
int sam_read1(samFile *fp, bam_hdr_t *h, bam1_t *b) { (void)fp; (void)h; (void)b; DNB_UNREACHABLE("sam_read1"); }
int sam_write1(samFile *fp, const bam_hdr_t *h, const bam1_t *b) { (void)fp; (void)h; (void)b; DNB_UNREACHABLE("sam_write1"); }
htsFile *hts_open(const char *fn, const char *mode) { (void)fn; (void)mode; DNB_UNREACHABLE("hts_open"); }
int hts_close(htsFile *fp) { (void)fp; DNB_UNREACHABLE("hts_close"); }

/* pod5 C API: only init/terminate are referenced (src/detect.cpp:816,917) */
int pod5_init(void) { return 0; }
int pod5_terminate(void) { return 0; }
