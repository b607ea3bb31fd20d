/*
 * rav1d_b200 host layer: a decoder loop over many streams in one process (one CUDA context), for the conformance tests.
 * Same calls as the reference's CLI (tools/dav1d.c: input_open / dav1d_send_data / dav1d_get_picture / md5 muxer with
 * output_verify, tools/output/md5.c), linked against the same host layer as dav1d_b200.
 *
 * usage: dav1d_b200_multi <list>      list: one "<stream path> <expected md5> [<apply_grain 0|1>]" per line (film grain
 *                                     defaults to 0: the reference's CLI switches it off under the md5 muxer unless
 *                                     --filmgrain is given, tools/dav1d_cli_parse.c:420-426)
 * prints one line per stream: "ok <path>", "MISMATCH <path>", "UNSUPPORTED <path>: why" or "ERROR <path>: what";
 * exit status 0 iff every stream was ok.
 */
#include "config.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dav1d/dav1d.h"
#include "input/input.h"
#include "output/output.h"

#include "host_frame.h"

static int decode_one(const char *const path, const char *const md5, const int apply_grain, const char **const what) {
    DemuxerContext *in = NULL;
    MuxerContext *out = NULL;
    Dav1dContext *c = NULL;
    unsigned fps[2], total, timebase[2];
    *what = "";
    if (input_open(&in, NULL, path, fps, &total, timebase) < 0) { *what = "cannot open"; return -1; }
    Dav1dSettings s;
    dav1d_default_settings(&s);
    s.n_threads = 2;
    s.max_frame_delay = 2;
    s.apply_grain = apply_grain;
    s.strict_std_compliance = 1;      /* as the CLI (tools/dav1d_cli_parse.c:293) */
    if (dav1d_open(&c, &s)) { input_close(in); *what = "dav1d_open"; return -1; }
    Dav1dData data;
    memset(&data, 0, sizeof(data));
    int res = input_read(in, &data), err = 0;
    if (res < 0) { *what = "empty input"; err = 1; }
    while (!err) {
        Dav1dPicture p;
        memset(&p, 0, sizeof(p));
        if (data.sz > 0) {
            res = dav1d_send_data(c, &data);
            if (res < 0 && res != DAV1D_ERR(EAGAIN)) { dav1d_data_unref(&data); *what = "dav1d_send_data"; err = 1; break; }
        }
        res = dav1d_get_picture(c, &p);
        if (res < 0) {
            if (res != DAV1D_ERR(EAGAIN)) { *what = "dav1d_get_picture"; err = 1; break; }
        } else {
            if (!out && output_open(&out, getenv("RB200_MULTI_YUV") ? "yuv" : "md5", getenv("RB200_MULTI_YUV") ? getenv("RB200_MULTI_YUV") : getenv("RB200_MULTI_PRINT") ? "-" : "/dev/null", &p.p, fps) < 0) { *what = "output_open"; err = 1; dav1d_picture_unref(&p); break; }
            if (output_write(out, &p) < 0) { *what = "output_write"; err = 1; break; }   /* unrefs p */
        }
        if (data.sz == 0 && input_read(in, &data) < 0) break;
    }
    if (data.sz > 0) dav1d_data_unref(&data);
    while (!err) {   /* drain */
        Dav1dPicture p;
        memset(&p, 0, sizeof(p));
        res = dav1d_get_picture(c, &p);
        if (res < 0) { if (res != DAV1D_ERR(EAGAIN)) { *what = "dav1d_get_picture (drain)"; err = 1; } break; }
        if (!out && output_open(&out, getenv("RB200_MULTI_YUV") ? "yuv" : "md5", getenv("RB200_MULTI_YUV") ? getenv("RB200_MULTI_YUV") : getenv("RB200_MULTI_PRINT") ? "-" : "/dev/null", &p.p, fps) < 0) { *what = "output_open"; err = 1; dav1d_picture_unref(&p); break; }
        if (output_write(out, &p) < 0) { *what = "output_write"; err = 1; break; }
    }
    input_close(in);
    dav1d_close(&c);
    int r = err ? -1 : 0;
    if (out) {
        if (err || getenv("RB200_MULTI_PRINT") || getenv("RB200_MULTI_YUV")) output_close(out);     /* RB200_MULTI_PRINT: the md5 muxer prints the hash to stdout */
        else if (output_verify(out, md5)) r = 1;
    } else if (!err) { *what = "no picture"; r = -1; }
    return r;
}

int main(int argc, char **argv) {
    if (argc != 2) { fprintf(stderr, "usage: %s <list of \"path md5\" lines>\n", argv[0]); return 2; }
    FILE *const l = fopen(argv[1], "r");
    if (!l) { perror(argv[1]); return 2; }
    const char *why;
    rb200_host_unsupported(&why);   /* report instead of exiting */
    char line[4300], path[4096], md5[64];
    int bad = 0, n = 0;
    while (fgets(line, sizeof(line), l)) {
        int grain = 0;
        if (sscanf(line, "%4095s %63s %d", path, md5, &grain) < 2) continue;
        const char *what;
        const int r = decode_one(path, md5, grain, &what);
        n++;
        if (rb200_host_unsupported(&why)) { printf("UNSUPPORTED %s: %s\n", path, why); bad++; }
        else if (r < 0) { printf("ERROR %s: %s\n", path, what); bad++; }
        else if (r > 0) { printf("MISMATCH %s\n", path); bad++; }
        else printf("ok %s\n", path);
        fflush(stdout);
    }
    fclose(l);
    fprintf(stderr, "%d streams, %d not ok\n", n, bad);
    return bad ? 1 : 0;
}
