// preprocess.cu — image preprocessing on the device: what the reference does per image on the host with
// PIL + numpy (src/mm_utils.py:152-182: expand2square, then CLIPImageProcessor.preprocess of transformers 4.46.3 = PIL bicubic
// resize of the uint8 image to the shortest edge, center crop) as two batched integer kernels over a whole batch of decoded
// images of arbitrary sizes.  Output: uint8 pixels [B, 3, S, S] at the tower's resolution, which setok_vit_forward_u8
// rescales + normalises inside its patch-embedding pass.
//
// The arithmetic is Pillow's ImagingResample for 8-bit images restated: separable convolution with per-output-pixel taps in
// 22-bit fixed point (the tables are built by the host in double precision exactly as Resample.c:precompute_coeffs does, so
// there is no device-side float arithmetic to disagree about), horizontal pass then vertical pass, each rounding to uint8
// with clipping.  Integer work, bit-exact against PIL.  Padding to a square is a virtual canvas (no copy) and the center crop
// restricts both passes to the pixels that survive, so a 4000 x 3000 photo costs 224 columns per source row, not 4000.
//
// HBM-bound byte work: kernel 1 reads each source row once (coalesced 96-byte runs of HWC pixels, taps from L1) and writes
// the [rows, S, 3] intermediate; kernel 2 reads that intermediate column-coalesced and writes planar rows.
#include "common.cuh"

namespace setok {
namespace {

constexpr int PP_BITS = 22;

__device__ __forceinline__ uint8_t pp_clip8(int v) {
  v >>= PP_BITS;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[b][y - y0][x'][c] = clip8(sum_k canvas[y][xmin + k][c] * kx[x'][k]),  x' in the crop window, y in [y0, y1)
__global__ void __launch_bounds__(128) resize_h_kernel(const setok_resize_desc* __restrict__ descs, const int32_t* __restrict__ tables,
                                                       uint8_t* __restrict__ workspace, int S, uchar3 bg) {
  const setok_resize_desc d = descs[blockIdx.z];
  const int y = d.y0 + static_cast<int>(blockIdx.y);
  if (y >= d.y1) return;
  const int sy = y - d.pad_y;                              // source row of canvas row y (outside [0, H): background)
  const bool row_in = sy >= 0 && sy < d.H;
  const uint8_t* srow = d.src + static_cast<long long>(row_in ? sy : 0) * d.W * 3;
  uint8_t* trow = workspace + d.tmp_off + static_cast<long long>(y - d.y0) * S * 3;
  for (int xo = blockIdx.x * blockDim.x + threadIdx.x; xo < S; xo += gridDim.x * blockDim.x) {
    const int32_t* tab = tables + d.kx_off + static_cast<long long>(xo) * (2 + d.ksize_x);
    uchar3 o;
    if (d.identity_x) {                                    // PIL skips a pass whose size does not change
      const int cx = d.left + xo, sx = cx - d.pad_x;
      const bool in = row_in && sx >= 0 && sx < d.W;
      o = in ? make_uchar3(srow[sx * 3], srow[sx * 3 + 1], srow[sx * 3 + 2]) : bg;
    } else {
      const int xmin = tab[0], n = tab[1];
      int a0 = 1 << (PP_BITS - 1), a1 = a0, a2 = a0;
      for (int k = 0; k < n; ++k) {
        const int sx = xmin + k - d.pad_x;
        const int w = tab[2 + k];
        const bool in = row_in && sx >= 0 && sx < d.W;
        const int p0 = in ? srow[sx * 3] : bg.x, p1 = in ? srow[sx * 3 + 1] : bg.y, p2 = in ? srow[sx * 3 + 2] : bg.z;
        a0 += p0 * w; a1 += p1 * w; a2 += p2 * w;
      }
      o = make_uchar3(pp_clip8(a0), pp_clip8(a1), pp_clip8(a2));
    }
    trow[xo * 3] = o.x; trow[xo * 3 + 1] = o.y; trow[xo * 3 + 2] = o.z;
  }
}

// out[b][c][y'][x'] = clip8(sum_k tmp[ymin + k - y0][x'][c] * ky[y'][k])
__global__ void __launch_bounds__(128) resize_v_kernel(const setok_resize_desc* __restrict__ descs, const int32_t* __restrict__ tables,
                                                       const uint8_t* __restrict__ workspace, uint8_t* __restrict__ out, int S) {
  const setok_resize_desc d = descs[blockIdx.z];
  const int yo = blockIdx.y;
  const uint8_t* tmp = workspace + d.tmp_off;
  const int32_t* tab = tables + d.ky_off + static_cast<long long>(yo) * (2 + d.ksize_y);
  uint8_t* obase = out + static_cast<long long>(blockIdx.z) * 3 * S * S + static_cast<long long>(yo) * S;
  for (int xo = blockIdx.x * blockDim.x + threadIdx.x; xo < S; xo += gridDim.x * blockDim.x) {
    uchar3 o;
    if (d.identity_y) {
      const uint8_t* t = tmp + (static_cast<long long>(d.top + yo - d.y0) * S + xo) * 3;
      o = make_uchar3(t[0], t[1], t[2]);
    } else {
      const int ymin = tab[0], n = tab[1];
      int a0 = 1 << (PP_BITS - 1), a1 = a0, a2 = a0;
      for (int k = 0; k < n; ++k) {
        const uint8_t* t = tmp + (static_cast<long long>(ymin + k - d.y0) * S + xo) * 3;
        const int w = tab[2 + k];
        a0 += t[0] * w; a1 += t[1] * w; a2 += t[2] * w;
      }
      o = make_uchar3(pp_clip8(a0), pp_clip8(a1), pp_clip8(a2));
    }
    obase[xo] = o.x;
    obase[static_cast<long long>(S) * S + xo] = o.y;
    obase[2LL * S * S + xo] = o.z;
  }
}

}  // namespace
}  // namespace setok

using namespace setok;

extern "C" int setok_preprocess_u8(const setok_resize_desc* descs_dev, int B, int max_rows, const int32_t* tables_dev, int S,
                                   const uint8_t* background, uint8_t* out, void* workspace, size_t workspace_bytes, size_t workspace_needed,
                                   setok_stream_t stream_) {
  SETOK_NVTX("setok f3 image preprocessing");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SETOK_REQUIRE(descs_dev && tables_dev && out && workspace && background, SETOK_ERR_BAD_ARG, "preprocess_u8: null pointer");
  SETOK_REQUIRE(B > 0 && S > 0 && max_rows > 0, SETOK_ERR_BAD_ARG, "preprocess_u8: B=%d S=%d max_rows=%d", B, S, max_rows);
  SETOK_REQUIRE(B <= 65535 && max_rows <= 65535 && S <= 65535, SETOK_ERR_UNSUPPORTED, "preprocess_u8: grid limits exceeded");
  SETOK_REQUIRE(workspace_bytes >= workspace_needed, SETOK_ERR_WORKSPACE, "preprocess_u8: workspace too small");
  const uchar3 bg = make_uchar3(background[0], background[1], background[2]);
  const int gx = ceil_div(S, 128);
  resize_h_kernel<<<dim3(gx, max_rows, B), 128, 0, stream>>>(descs_dev, tables_dev, static_cast<uint8_t*>(workspace), S, bg);
  SETOK_LAUNCH_CHECK();
  resize_v_kernel<<<dim3(gx, S, B), 128, 0, stream>>>(descs_dev, tables_dev, static_cast<const uint8_t*>(workspace), out, S);
  SETOK_LAUNCH_CHECK();
  return SETOK_OK;
}
