// Overlay drawing on frames that are resident in HBM: the pixels cv2.rectangle / cv2.line /
// cv2.putText would have written for reference utils/helpers.py:126-179 (draw_bbox, draw_bbox_info), called per face at
// reference main.py:144-148.  Byte work, HBM-/latency-bound: one CTA per frame walks the frame's draw list.
//
// The host side (scrfd_arcface_facerecognition_b200/overlay.py) lowers every cv2 call of the two reference functions to
// inclusive rectangles and 1-bit mask blits:
//   * cv2.rectangle(..., 1)        = four one-pixel rows / columns between the corner coordinates
//   * cv2.rectangle(..., FILLED)   = one rectangle
//   * cv2.line(..., thickness 3) on an axis-aligned segment = the band two pixels either side of the segment plus a
//     radius-2 disc (rows of 1, 3, 5, 3, 1 pixels) at both end points -- seven rectangles
//   * cv2.putText                  = the glyph coverage cv2 itself rendered once per label, blitted as a mask
// Everything one face draws has one colour, so the order of its commands does not matter; faces overlap, so the
// order BETWEEN faces does: commands are grouped per face and the CTA puts a barrier between groups.
#include "b2f_common.cuh"
#include "../../include/b2f.h"

#include <atomic>

namespace b2f {
extern std::atomic<long long> g_launches;

__global__ void __launch_bounds__(256)
draw_overlay_kernel(uint8_t* __restrict__ frames, int h, int w, const b2f_draw_cmd* __restrict__ cmds,
                    const int* __restrict__ frame_groups, const int* __restrict__ group_cmds,
                    const uint8_t* __restrict__ masks) {
  const int f = blockIdx.x;
  const int g0 = frame_groups[f], g1 = frame_groups[f + 1];
  if (g0 == g1) return;
  uint8_t* img = frames + (size_t)f * h * w * 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int g = g0; g < g1; ++g) {
    const int c0 = group_cmds[g], c1 = group_cmds[g + 1];
    for (int c = c0 + warp; c < c1; c += n_warps) {               // a warp per command, lanes over its pixels
      const b2f_draw_cmd cmd = cmds[c];
      const uint8_t b = (uint8_t)(cmd.bgr & 0xFF), gr = (uint8_t)((cmd.bgr >> 8) & 0xFF), r = (uint8_t)((cmd.bgr >> 16) & 0xFF);
      if (cmd.kind == 0) {
        const int x0 = max(cmd.x0, 0), y0 = max(cmd.y0, 0), x1 = min(cmd.x1, w - 1), y1 = min(cmd.y1, h - 1);
        if (x0 > x1 || y0 > y1) continue;
        const int rw = x1 - x0 + 1, area = rw * (y1 - y0 + 1);
        for (int i = lane; i < area; i += 32) {
          const int y = y0 + i / rw, x = x0 + i % rw;
          uint8_t* p = img + ((size_t)y * w + x) * 3;
          p[0] = b, p[1] = gr, p[2] = r;
        }
      } else {
        // mask blit: x0, y0 = where mask pixel (0, 0) lands; x1, y1 = mask width, height
        const uint8_t* m = masks + cmd.mask_off;
        const int mw = cmd.x1, area = cmd.x1 * cmd.y1;
        for (int i = lane; i < area; i += 32) {
          if (!m[i]) continue;
          const int y = cmd.y0 + i / mw, x = cmd.x0 + i % mw;
          if (x < 0 || y < 0 || x >= w || y >= h) continue;
          uint8_t* p = img + ((size_t)y * w + x) * 3;
          p[0] = b, p[1] = gr, p[2] = r;
        }
      }
    }
    __syncthreads();                                             // the next face may paint over this one
  }
}

}  // namespace b2f

extern "C" int b2f_draw_overlay(uint8_t* frames, int batch, int h, int w, const b2f_draw_cmd* cmds,
                                const int* frame_groups, const int* group_cmds, const uint8_t* masks, void* stream) {
  B2F_REQUIRE(frames != nullptr && batch >= 0 && h > 0 && w > 0, "b2f_draw_overlay: bad frame batch %d x %d x %d", batch, h, w);
  B2F_REQUIRE(frame_groups != nullptr && group_cmds != nullptr, "b2f_draw_overlay: draw list offsets missing");
  if (batch == 0) return 0;
  b2f::draw_overlay_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(frames, h, w, cmds, frame_groups, group_cmds, masks);
  b2f::g_launches.fetch_add(1);
  B2F_LAUNCH_CHECK();
  return 0;
}
