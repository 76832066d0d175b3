/*
 * TEST INFRASTRUCTURE ONLY (oracle/): link-time stubs for the 14 htslib entry points
 * that the reference's EM.cpp pulls in through BamWriter.h / SamHeader.cpp
 * (BamWriter.h:39-146, SamHeader.cpp).  The oracle build (oracle/Makefile) compiles
 * the reference sources where they lie under /root/reference and never runs the
 * reference's own build system, so libhts.a is not available.  None of these are
 * reachable unless rsem-run-em is given "-b", which the oracle harness never does
 * (posterior BAM output is SURVEY.md section 8(f) "next", host I/O only).
 * Every stub aborts loudly if it is ever called.
 */
#include <stdio.h>
#include <stdlib.h>

#define HTS_STUB(name)                                                            \
    void *name(void) {                                                            \
        fprintf(stderr, "oracle/_ref: htslib stub '%s' called (-b unsupported)\n", \
                #name);                                                           \
        abort();                                                                  \
        return NULL;                                                              \
    }

HTS_STUB(bam_aux_append)
HTS_STUB(bam_aux_get)
HTS_STUB(bam_destroy1)
HTS_STUB(bam_hdr_destroy)
HTS_STUB(bam_init1)
HTS_STUB(hts_close)
HTS_STUB(hts_open)
HTS_STUB(hts_set_fai_filename)
HTS_STUB(hts_set_threads)
HTS_STUB(sam_hdr_parse)
HTS_STUB(sam_hdr_read)
HTS_STUB(sam_hdr_write)
HTS_STUB(sam_read1)
HTS_STUB(sam_write1)
