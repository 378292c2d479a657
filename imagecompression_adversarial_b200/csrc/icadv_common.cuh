// Shared host-side helpers: error reporting across the C ABI, geometry of the two contraction forms.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/icadv.h"

namespace icadv {

void set_error(const char* fmt, ...);

#define ICADV_CUDA_TRY(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      icadv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ICADV_ECUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define ICADV_REQUIRE(cond, ...)       \
  do {                                 \
    if (!(cond)) {                     \
      icadv::set_error(__VA_ARGS__);   \
      return ICADV_EINVAL;             \
    }                                  \
  } while (0)

inline cudaStream_t as_stream(icadv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// One tap of a contraction launch, in "tile space" (see icadv.h):
//   SCONV: tile space = output pixels; input pixel = stride*(i,j) + (kh-p, kw-p).  For stride 2 the
//          input is addressed through 4 parity planes: plane = (a,b), plane coords (i+dy, j+dx).
//   TCONV: tile space = input pixels, one launch per output parity (a,b); input pixel (i+dy, j+dx).
struct Tap {
  int16_t plane;  // input parity plane index a*2+b (SCONV stride 2), else 0
  int16_t dx, dy; // offset in plane / input coordinates
  int16_t wtap;   // kh*ksize + kw : row block of the packed weight
};

constexpr int kMaxTaps = 32;

struct Geometry {
  int form, ksize, stride, pad;
  int in_h, in_w, out_h, out_w;
  int tile_h, tile_w;  // extent of tile space
  int n_launch;        // 1 (SCONV) or stride*stride (TCONV)
  int n_taps[4];
  Tap taps[4][kMaxTaps];
  int out_a[4], out_b[4];  // output parity of each launch (TCONV)
};

// Fills g from the descriptor; returns 0 or ICADV_EINVAL.
int make_geometry(const icadv_conv_desc* d, Geometry* g);

// tcgen05 weight gradient (icadv_wgrad_tc.cu): eligibility and launch; dwpack [taps][n_ch][k_ch] is overwritten
int wgrad_tc_supported(const icadv_conv_desc* d);
int wgrad_tc(const icadv_conv_desc* d, const Geometry& g, const float* gout, float* dwpack, cudaStream_t stream);

}  // namespace icadv
