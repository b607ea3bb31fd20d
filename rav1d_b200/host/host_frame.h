/* rav1d_b200 host layer: what recon_batch_tmpl.c needs from the frame life cycle (host_frame.c). */
#ifndef RB200_HOST_FRAME_H
#define RB200_HOST_FRAME_H

#include "rb200_host.h"

struct Dav1dFrameContext;

/* The batch of the frame `f` is decoding.  Blocks of different tiles are appended from different task threads, so the
 * pair brackets one block's records.  The first call of a frame also waits until every reference picture of the frame
 * has been handed to the backend, which keeps the hand-over in decode order. */
RbHostBatch *rb_host_block_begin(const struct Dav1dFrameContext *f);
void rb_host_block_end(const struct Dav1dFrameContext *f);

/* pal_pred over bw4 x bh4 units of plane `pl`: a record { palette padded to 16 bytes, w * h index bytes } in the
 * batch's palette buffer and a wavefront item of mode 14 that names it. */
void rb_host_push_palette(RbHostBatch *b, int x4, int y4, int w4_end, int h4_end, int pl, int bw4, int bh4,
                          const void *pal, int pal_bytes, const uint8_t *idx);

/* Whether the last decoded stream needed a record the batch formats do not have (and what).  Calling it also turns off
 * the default behaviour of leaving the process with status 3 in that case. */
int rb200_host_unsupported(const char **why);

#endif
