// Column-strip walking shared by the persistent strip kernels (espcn_fused.cuh, conv_strip.cu): a CTA owns a contiguous range of
// "units" (one unit = one image row of one column strip); the range is cut at strip boundaries into segments (frame n, strip s,
// rows [ya, ya + len)), each `apron` virtual rows longer than its unit count (the vertical receptive field a segment has to
// warm up).  One thread tabulates the segments in shared memory at kernel start (the only divisions of the kernel); every role
// then follows the table with a cursor held in registers.
#pragma once
#include <cstdint>

namespace srk {

constexpr int kEfMaxSegs = 64;            // strip segments one CTA can follow; the host splits a call so that no CTA sees more
constexpr int kEfTabInts = 72 + 3 * 64;   // segment table: v0[0..64] (v0[nseg] = V), nseg at [71], then n[64], s[64], ya[64]

struct EfSeg {
  int i, v0, v1, n, s, ya;  // segment index, its virtual rows [v0, v1), frame, strip, first image row
};
__device__ __forceinline__ void ef_seg_load(EfSeg& c, const int* tab, int i) {
  c.i = i;
  c.v0 = tab[i];
  c.v1 = tab[i + 1];
  c.n = tab[72 + i];
  c.s = tab[72 + 64 + i];
  c.ya = tab[72 + 128 + i];
}
// position the cursor on virtual row v (< V; v never decreases)
__device__ __forceinline__ void ef_seg_seek(EfSeg& c, const int* tab, int v) {
  while (v >= c.v1) ef_seg_load(c, tab, c.i + 1);
}
static __device__ __noinline__ void ef_build_segments(int* tab, uint32_t u0, uint32_t u1, int hb, int y0, int strips, int apron) {
  int v = 0, i = 0;
  while (u0 < u1 && i < kEfMaxSegs) {
    const uint32_t qq = u0 / uint32_t(hb);
    const uint32_t off = u0 - qq * uint32_t(hb);
    uint32_t len = uint32_t(hb) - off;
    if (len > u1 - u0) len = u1 - u0;
    const uint32_t n = qq / uint32_t(strips);
    tab[i] = v;
    tab[72 + i] = int(n);
    tab[72 + 64 + i] = int(qq - n * uint32_t(strips));
    tab[72 + 128 + i] = y0 + int(off);
    v += int(len) + apron;
    u0 += len;
    ++i;
  }
  tab[i] = v;
  tab[71] = i;
}

}  // namespace srk
