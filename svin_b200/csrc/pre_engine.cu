// Image pre-processing in front of the detector (sm_100a): resize -> 3x3 median -> CLAHE / equalizeHist.
//
// Reference path: Subscriber::imageCallback (okvis_ros/src/Subscriber.cpp:123-147) calls cv::resize, cv::medianBlur,
// cv::CLAHE::apply / cv::equalizeHist on every incoming image.  The arithmetic restated here is OpenCV 4.x's
// (modules/imgproc/src/resize.cpp, median_blur.cpp, histogram.cpp, clahe.cpp) for 8-bit single-channel images;
// results are bit-exact with cv2 4.13 (tests/golden/preprocess_golden.npz) - integer arithmetic for the resize and the
// histograms, correctly rounded single-precision operations (no FMA contraction) for the LUT scale and the CLAHE
// interpolation.
//
// Byte work, HBM-bound in principle: per image of S source and D output pixels the chain reads
// S + 3 D and writes up to 3 D bytes; a batch of images is processed by the same launches (blockIdx.y = image).
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.hpp"
#include "host_pool.hpp"

namespace svin {
namespace {

struct PreDev {
  int sw, sh, spitch;   // source size, device pitch of the source rows
  int w, h;             // output size
  int mode;             // 0 copy, 1 exact 2x decimation (area fast path), 2 fixed-point bilinear
  int tiles, tw, th;    // CLAHE grid, tile size (of the padded image)
  int method;
  int clip;             // integer clip limit (0 = none)
  float lut_scale;      // 255 / tile area
  const uint8_t* src;   // [n][sh][spitch]
  uint8_t *a, *b;       // [n][h][w] ping-pong stages
  const int* xofs;      // bilinear taps: [w] source column, [w][2] weights (2^11 fixed point)
  const short* xw;
  const int* yofs;      // [h][2] clipped source rows, [h][2] weights
  const short* yw;
  unsigned* hist;       // [n][tiles*tiles][256]
  uint8_t* lut;         // [n][tiles*tiles][256]
};

// ---- resize ---------------------------------------------------------------------------------------------------
// resizeAreaFast_Invoker with scale 2: (a + b + c + d + 2) >> 2; blocks cut by the border average what exists.
__device__ __forceinline__ uint8_t decimate_px(const PreDev& p, const uint8_t* s, int x, int y) {
  const int x0 = 2 * x, y0 = 2 * y;
  int sum = 0, cnt = 0;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
      if (y0 + dy < p.sh && x0 + dx < p.sw) {
        sum += s[(size_t)(y0 + dy) * p.spitch + x0 + dx];
        ++cnt;
      }
  int v = 0;
  if (cnt == 4)
    v = (sum + 2) >> 2;
  else if (cnt > 0)
    v = __float2int_rn(__fdiv_rn((float)sum, (float)cnt));  // saturate_cast<uchar>((float)sum / count)
  return (uint8_t)min(max(v, 0), 255);
}
// Four output pixels per thread: two 8-byte row segments in, one 4-byte store out (the source pitch is a multiple
// of 16 and the stage buffers are dense, so the vector accesses are aligned whenever the output width is a multiple
// of 4; other widths and the border blocks of odd sources take the scalar path).
__global__ void k_pre_decimate2(PreDev p) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, img = blockIdx.z;
  if (x4 >= p.w) return;
  const uint8_t* s = p.src + ((size_t)img * p.sh) * p.spitch;
  uint8_t* d = p.a + ((size_t)img * p.h + y) * p.w;
  if ((p.w & 3) == 0 && 2 * x4 + 8 <= p.sw && 2 * y + 2 <= p.sh) {
    const uint2 r0 = *reinterpret_cast<const uint2*>(s + (size_t)(2 * y) * p.spitch + 2 * x4);
    const uint2 r1 = *reinterpret_cast<const uint2*>(s + (size_t)(2 * y + 1) * p.spitch + 2 * x4);
    auto quad = [](unsigned a, unsigned b, int k) {  // pixels 2k, 2k+1 of both rows
      const unsigned s0 = (a >> (16 * k)) & 0xffffu, s1 = (b >> (16 * k)) & 0xffffu;
      return ((s0 & 255u) + (s0 >> 8) + (s1 & 255u) + (s1 >> 8) + 2u) >> 2;
    };
    const unsigned o = quad(r0.x, r1.x, 0) | (quad(r0.x, r1.x, 1) << 8) | (quad(r0.y, r1.y, 0) << 16) |
                       (quad(r0.y, r1.y, 1) << 24);
    *reinterpret_cast<unsigned*>(d + x4) = o;
  } else {
    for (int x = x4; x < min(x4 + 4, p.w); ++x) d[x] = decimate_px(p, s, x, y);
  }
}
// HResizeLinear<uchar,int,short,2048> + VResizeLinear<uchar,int,short>: taps from the host tables.
__global__ void k_pre_bilinear(PreDev p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
  if (x >= p.w) return;
  const uint8_t* s = p.src + ((size_t)img * p.sh) * p.spitch;
  const int sx = p.xofs[x], sx1 = min(sx + 1, p.sw - 1);
  const int a0 = p.xw[2 * x], a1 = p.xw[2 * x + 1];
  const int r0 = p.yofs[2 * y], r1 = p.yofs[2 * y + 1];
  const int b0 = p.yw[2 * y], b1 = p.yw[2 * y + 1];
  const int h0 = s[(size_t)r0 * p.spitch + sx] * a0 + s[(size_t)r0 * p.spitch + sx1] * a1;
  const int h1 = s[(size_t)r1 * p.spitch + sx] * a0 + s[(size_t)r1 * p.spitch + sx1] * a1;
  const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
  p.a[((size_t)img * p.h + y) * p.w + x] = (uint8_t)min(max(v, 0), 255);
}
__global__ void k_pre_copy(PreDev p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
  if (x >= p.w) return;
  p.a[((size_t)img * p.h + y) * p.w + x] = p.src[((size_t)img * p.sh + y) * p.spitch + x];
}

// ---- 3x3 median, BORDER_REPLICATE (median_blur.cpp) -----------------------------------------------------------
__device__ __forceinline__ void srt(int& a, int& b) {
  const int lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}
__global__ void k_pre_median3(PreDev p, const uint8_t* in, uint8_t* out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
  if (x >= p.w) return;
  const uint8_t* s = in + (size_t)img * p.h * p.w;
  int v[9];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = min(max(y + dy, 0), p.h - 1);
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) v[(dy + 1) * 3 + dx + 1] = s[(size_t)yy * p.w + min(max(x + dx, 0), p.w - 1)];
  }
  // 19-exchange median-of-9 network
  srt(v[1], v[2]); srt(v[4], v[5]); srt(v[7], v[8]); srt(v[0], v[1]); srt(v[3], v[4]); srt(v[6], v[7]);
  srt(v[1], v[2]); srt(v[4], v[5]); srt(v[7], v[8]); srt(v[0], v[3]); srt(v[5], v[8]); srt(v[4], v[7]);
  srt(v[3], v[6]); srt(v[1], v[4]); srt(v[2], v[5]); srt(v[4], v[7]); srt(v[4], v[2]); srt(v[6], v[4]);
  srt(v[4], v[2]);
  out[((size_t)img * p.h + y) * p.w + x] = (uint8_t)v[4];
}

// ---- histograms: one per CLAHE tile (of the image padded with BORDER_REFLECT_101), or one per image ------------
constexpr int kHistRows = 32;  // rows of a tile per CTA
__global__ void __launch_bounds__(256) k_pre_hist(PreDev p, const uint8_t* in) {
  __shared__ unsigned sh[256];
  const int tile = blockIdx.x, stripe = blockIdx.y, img = blockIdx.z;
  const int T = p.method == SVIN_HIST_CLAHE ? p.tiles : 1;
  const int tw = p.method == SVIN_HIST_CLAHE ? p.tw : p.w, th = p.method == SVIN_HIST_CLAHE ? p.th : p.h;
  const int tx = tile % T, ty = tile / T;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint8_t* s = in + (size_t)img * p.h * p.w;
  const int y_begin = stripe * kHistRows, y_end = min(y_begin + kHistRows, th);
  for (int yy = y_begin; yy < y_end; ++yy) {
    int y = ty * th + yy;
    if (y >= p.h) y = 2 * (p.h - 1) - y;  // BORDER_REFLECT_101
    const uint8_t* r = s + (size_t)y * p.w;
    const int xb = tx * tw;
    if (((p.w | xb | tw) & 3) == 0 && xb + tw <= p.w) {   // aligned 4-byte loads, no reflected columns
      for (int xx = 4 * threadIdx.x; xx < tw; xx += 4 * 256) {
        const unsigned v4 = *reinterpret_cast<const unsigned*>(r + xb + xx);
        atomicAdd(&sh[v4 & 255u], 1u);
        atomicAdd(&sh[(v4 >> 8) & 255u], 1u);
        atomicAdd(&sh[(v4 >> 16) & 255u], 1u);
        atomicAdd(&sh[v4 >> 24], 1u);
      }
    } else {
      for (int xx = threadIdx.x; xx < tw; xx += 256) {
        int x = xb + xx;
        if (x >= p.w) x = 2 * (p.w - 1) - x;
        atomicAdd(&sh[r[x]], 1u);
      }
    }
  }
  __syncthreads();
  const unsigned v = sh[threadIdx.x];
  if (v) atomicAdd(&p.hist[((size_t)img * T * T + tile) * 256 + threadIdx.x], v);
}

// ---- LUT per tile / image (256 threads) --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pre_lut(PreDev p) {
  __shared__ int sc[256];
  __shared__ int red[8];
  const int tile = blockIdx.x, img = blockIdx.y, t = threadIdx.x;
  const int T = p.method == SVIN_HIST_CLAHE ? p.tiles : 1;
  const size_t base = ((size_t)img * T * T + tile) * 256;
  int hv = (int)p.hist[base + t];
  const int lane = t & 31, wid = t >> 5;
  if (p.method == SVIN_HIST_CLAHE) {
    // CLAHE_CalcLut_Body: clip, redistribute the excess (batch + one extra for every residualStep-th bin)
    if (p.clip > 0) {
      int ex = max(hv - p.clip, 0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ex += __shfl_xor_sync(0xffffffffu, ex, o);
      if (lane == 0) red[wid] = ex;
      __syncthreads();
      int clipped = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) clipped += red[k];
      hv = min(hv, p.clip);
      const int batch = clipped / 256;
      int residual = clipped - batch * 256;
      hv += batch;
      if (residual != 0) {
        const int step = max(256 / residual, 1);
        if (t % step == 0 && t / step < residual) hv += 1;
      }
    }
  }
  // inclusive scan of the 256 bins
  sc[t] = hv;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const int add = t >= o ? sc[t - o] : 0;
    __syncthreads();
    sc[t] += add;
    __syncthreads();
  }
  uint8_t out;
  if (p.method == SVIN_HIST_CLAHE) {
    out = (uint8_t)min(max(__float2int_rn(__fmul_rn((float)sc[t], p.lut_scale)), 0), 255);
  } else {
    // cv::equalizeHist: first non-empty bin i maps to 0, lut[j > i] = round((cdf[j] - hist[i]) * 255 / (total - hist[i]))
    const unsigned nz = __ballot_sync(0xffffffffu, hv != 0);
    if (lane == 0) red[wid] = nz ? wid * 32 + __ffs(nz) - 1 : 256;
    __syncthreads();
    int first = 256;
#pragma unroll
    for (int k = 0; k < 8; ++k) first = min(first, red[k]);
    const int total = p.w * p.h;
    const int h0 = first < 256 ? sc[first] - (first ? sc[first - 1] : 0) : 0;
    if (h0 == total) {
      out = (uint8_t)first;  // constant image: dst = that value
    } else if (t <= first) {
      out = 0;
    } else {
      const float scale = __fdiv_rn(255.0f, (float)(total - h0));
      out = (uint8_t)min(max(__float2int_rn(__fmul_rn((float)(sc[t] - sc[first]), scale)), 0), 255);
    }
    if (h0 == total && t != first) out = (uint8_t)first;
  }
  p.lut[base + t] = out;
}

// ---- apply -----------------------------------------------------------------------------------------------------
// One CTA per strip of kApplyRows rows of one image: the image's LUTs (tiles^2 x 256 bytes, <= 64 KB) are staged in
// shared memory once per strip, then every thread maps four pixels at a time (one 4-byte load and store when the
// width is a multiple of 4).
constexpr int kApplyRows = 8;
__global__ void __launch_bounds__(256) k_pre_apply(PreDev p, const uint8_t* in, uint8_t* out) {
  extern __shared__ uint8_t sL[];
  const int img = blockIdx.y, y0 = blockIdx.x * kApplyRows;
  const int T = p.method == SVIN_HIST_CLAHE ? p.tiles : 1;
  {
    const uint4* g = reinterpret_cast<const uint4*>(p.lut + (size_t)img * T * T * 256);
    uint4* d = reinterpret_cast<uint4*>(sL);
    for (int k = threadIdx.x; k < T * T * 16; k += 256) d[k] = g[k];
  }
  // per-column interpolation terms (tx1, tx2, xa) of CLAHE, computed once per strip
  float2* colT = reinterpret_cast<float2*>(sL + (size_t)T * T * 256);
  const float inv_tw = __fdiv_rn(1.0f, (float)max(p.tw, 1)), inv_th = __fdiv_rn(1.0f, (float)max(p.th, 1));
  if (p.method == SVIN_HIST_CLAHE) {
    for (int x = threadIdx.x; x < p.w; x += 256) {
      const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
      const int tx1 = (int)floorf(txf);
      const float xa = __fsub_rn(txf, (float)tx1);
      const int t1 = max(tx1, 0), t2 = min(tx1 + 1, T - 1);
      colT[x] = make_float2(xa, __int_as_float(t1 | (t2 << 8)));
    }
  }
  __syncthreads();
  const bool vec = (p.w & 3) == 0;
  const int quads = (p.w + 3) >> 2;
  for (int yy = 0; yy < kApplyRows; ++yy) {
    const int y = y0 + yy;
    if (y >= p.h) break;
    const size_t row = ((size_t)img * p.h + y) * p.w;
    const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf);
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    const int ty2 = min(ty1 + 1, T - 1);
    ty1 = max(ty1, 0);
    for (int q = threadIdx.x; q < quads; q += 256) {
      const int x4 = 4 * q, nx = min(4, p.w - x4);
      unsigned vin = 0;
      if (vec)
        vin = *reinterpret_cast<const unsigned*>(in + row + x4);
      else
        for (int k = 0; k < nx; ++k) vin |= (unsigned)in[row + x4 + k] << (8 * k);
      unsigned vout = 0;
      if (p.method == SVIN_HIST_EQUALIZE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) vout |= (unsigned)sL[(vin >> (8 * k)) & 255u] << (8 * k);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= nx) break;
          const float2 ct = colT[x4 + k];
          const int tt = __float_as_int(ct.y), tx1 = tt & 255, tx2 = tt >> 8;
          const float xa = ct.x, xa1 = __fsub_rn(1.0f, xa);
          const int v = (vin >> (8 * k)) & 255u;
          const float l11 = sL[(ty1 * T + tx1) * 256 + v], l12 = sL[(ty1 * T + tx2) * 256 + v];
          const float l21 = sL[(ty2 * T + tx1) * 256 + v], l22 = sL[(ty2 * T + tx2) * 256 + v];
          const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
          const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
          const int r = __float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
          vout |= (unsigned)min(max(r, 0), 255) << (8 * k);
        }
      }
      if (vec)
        *reinterpret_cast<unsigned*>(out + row + x4) = vout;
      else
        for (int k = 0; k < nx; ++k) out[row + x4 + k] = (uint8_t)(vout >> (8 * k));
    }
  }
}

// cvRound(float) for non-negative short weights
inline short sat_short(float v) {
  const long r = lrintf(v);  // round half to even (default rounding mode)
  return (short)std::min(32767l, std::max(-32768l, r));
}
// Taps of the fixed-point bilinear resize (resize.cpp: xofs/alpha, yofs/beta); volatile keeps the compiler from
// contracting (d + 0.5) * scale - 0.5 into a fused multiply-add, which OpenCV's build does not use here.
void bilinear_taps(int dn, int sn, double scale, bool clamp_weights, std::vector<int>& ofs, std::vector<short>& wgt) {
  ofs.resize(clamp_weights ? dn : 2 * dn);
  wgt.resize(2 * dn);
  for (int d = 0; d < dn; ++d) {
    volatile double t = (d + 0.5) * scale;
    volatile double u = t - 0.5;
    volatile float f = (float)u;
    int s = (int)std::floor(f);
    volatile float fr = f - (float)s;
    float fx = fr;
    if (clamp_weights) {
      if (s < 0) { fx = 0; s = 0; }
      if (s >= sn - 1) { fx = 0; s = sn - 1; }
      ofs[d] = s;
    } else {
      ofs[2 * d] = std::min(std::max(s, 0), sn - 1);
      ofs[2 * d + 1] = std::min(std::max(s + 1, 0), sn - 1);
    }
    volatile float w0 = (1.0f - fx) * 2048.0f, w1 = fx * 2048.0f;
    wgt[2 * d] = sat_short(w0);
    wgt[2 * d + 1] = sat_short(w1);
  }
}

}  // namespace
}  // namespace svin

using svin::set_error;

struct svin_pre_ctx {
  int device = 0;
  SvinPreOptions opt{};
  svin::PreDev p{};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8]{};
  uint8_t *d_src = nullptr, *d_a = nullptr, *d_b = nullptr, *h_src = nullptr, *h_dst = nullptr;
  int *d_xofs = nullptr, *d_yofs = nullptr;
  short *d_xw = nullptr, *d_yw = nullptr;
  unsigned* d_hist = nullptr;
  uint8_t* d_lut = nullptr;
  const uint8_t* d_out = nullptr;
  int n_images = 0;
  svin::HostPool* pool = nullptr;
  SvinPreTimings tm{};
};

extern "C" {

int svin_pre_create(int device, const SvinPreOptions* o, svin_pre_ctx** out) {
  if (!o || !out || o->src_width < 8 || o->src_height < 8 || o->max_images < 1 || !(o->resize_factor > 0.0) ||
      o->histogram_method < SVIN_HIST_NONE || o->histogram_method > SVIN_HIST_CLAHE ||
      (o->histogram_method == SVIN_HIST_CLAHE && (o->clahe_tiles < 1 || o->clahe_tiles > 16))) {
    set_error("svin_pre_create: invalid options");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_error("svin_pre_create: no CUDA device (the svin_b200 engine has no CPU fallback)");
    return SVIN_ERR_NO_DEVICE;
  }
  SVIN_CUDA(cudaSetDevice(device));
  svin_pre_ctx* c = new svin_pre_ctx;
  c->device = device;
  c->opt = *o;
  svin::PreDev& p = c->p;
  p.sw = o->src_width;
  p.sh = o->src_height;
  p.spitch = (p.sw + 15) & ~15;
  const double f = o->resize_factor;
  if (f == 1.0) {
    p.mode = 0;
    p.w = p.sw;
    p.h = p.sh;
  } else {
    p.w = (int)std::nearbyint(p.sw * f);  // saturate_cast<int>(ssize.width * inv_scale_x): round half to even
    p.h = (int)std::nearbyint(p.sh * f);
    const double scale = 1.0 / f;
    const int is = (int)std::nearbyint(scale);
    p.mode = (std::fabs(scale - is) < 2.220446049250313e-16 && is == 2) ? 1 : 2;
  }
  if (p.w < 1 || p.h < 1) {
    delete c;
    set_error("svin_pre_create: empty output");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  p.method = o->histogram_method;
  p.tiles = p.method == SVIN_HIST_CLAHE ? o->clahe_tiles : 1;
  if (p.method == SVIN_HIST_CLAHE) {
    const int T = p.tiles;
    // clahe.cpp: when either dimension is not a multiple of the grid, BOTH are extended by tiles - (dim % tiles)
    // (a whole extra tile row/column of reflected pixels for the dimension that did divide)
    const bool pad = (p.w % T != 0) || (p.h % T != 0);
    const int pw = pad ? p.w + (T - p.w % T) : p.w, ph = pad ? p.h + (T - p.h % T) : p.h;
    p.tw = pw / T;
    p.th = ph / T;
    if (pw - p.w >= p.w || ph - p.h >= p.h) {
      delete c;
      set_error("svin_pre_create: image smaller than the CLAHE tile grid");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
    const int area = p.tw * p.th;
    p.lut_scale = 255.0f / (float)area;
    p.clip = 0;
    if (o->clahe_clip_limit > 0.0) p.clip = std::max((int)(o->clahe_clip_limit * area / 256), 1);
  }
  const int M = o->max_images;
  SVIN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& e : c->ev) SVIN_CUDA(cudaEventCreate(&e));
  SVIN_CUDA(cudaMalloc(&c->d_src, (size_t)M * p.sh * p.spitch));
  SVIN_CUDA(cudaMalloc(&c->d_a, (size_t)M * p.h * p.w));
  SVIN_CUDA(cudaMalloc(&c->d_b, (size_t)M * p.h * p.w));
  SVIN_CUDA(cudaMallocHost(&c->h_src, (size_t)M * p.sh * p.spitch));
  SVIN_CUDA(cudaMallocHost(&c->h_dst, (size_t)M * p.h * p.w));
  SVIN_CUDA(cudaMalloc(&c->d_hist, sizeof(unsigned) * (size_t)M * p.tiles * p.tiles * 256));
  SVIN_CUDA(cudaMalloc(&c->d_lut, (size_t)M * p.tiles * p.tiles * 256));
  if (p.mode == 2) {
    std::vector<int> xo, yo;
    std::vector<short> xw, yw;
    svin::bilinear_taps(p.w, p.sw, 1.0 / f, true, xo, xw);
    svin::bilinear_taps(p.h, p.sh, 1.0 / f, false, yo, yw);
    SVIN_CUDA(cudaMalloc(&c->d_xofs, sizeof(int) * xo.size()));
    SVIN_CUDA(cudaMalloc(&c->d_yofs, sizeof(int) * yo.size()));
    SVIN_CUDA(cudaMalloc(&c->d_xw, sizeof(short) * xw.size()));
    SVIN_CUDA(cudaMalloc(&c->d_yw, sizeof(short) * yw.size()));
    SVIN_CUDA(cudaMemcpy(c->d_xofs, xo.data(), sizeof(int) * xo.size(), cudaMemcpyHostToDevice));
    SVIN_CUDA(cudaMemcpy(c->d_yofs, yo.data(), sizeof(int) * yo.size(), cudaMemcpyHostToDevice));
    SVIN_CUDA(cudaMemcpy(c->d_xw, xw.data(), sizeof(short) * xw.size(), cudaMemcpyHostToDevice));
    SVIN_CUDA(cudaMemcpy(c->d_yw, yw.data(), sizeof(short) * yw.size(), cudaMemcpyHostToDevice));
  }
  p.src = c->d_src;
  p.a = c->d_a;
  p.b = c->d_b;
  p.xofs = c->d_xofs;
  p.xw = c->d_xw;
  p.yofs = c->d_yofs;
  p.yw = c->d_yw;
  p.hist = c->d_hist;
  p.lut = c->d_lut;
  if (p.tiles * p.tiles * 256 + 8 * p.w > 48 * 1024)
    SVIN_CUDA(cudaFuncSetAttribute(svin::k_pre_apply, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   p.tiles * p.tiles * 256 + 8 * p.w));
  c->pool = new svin::HostPool(std::max(0, std::min(16, (int)std::thread::hardware_concurrency()) - 1));
  *out = c;
  return SVIN_OK;
}

void svin_pre_destroy(svin_pre_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  delete c->pool;
  cudaFree(c->d_src); cudaFree(c->d_a); cudaFree(c->d_b); cudaFree(c->d_hist); cudaFree(c->d_lut);
  cudaFree(c->d_xofs); cudaFree(c->d_yofs); cudaFree(c->d_xw); cudaFree(c->d_yw);
  cudaFreeHost(c->h_src); cudaFreeHost(c->h_dst);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int svin_pre_output_size(svin_pre_ctx* c, int32_t* w, int32_t* h) {
  if (!c || !w || !h) {
    set_error("svin_pre_output_size: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  *w = c->p.w;
  *h = c->p.h;
  return SVIN_OK;
}

int svin_pre_upload(svin_pre_ctx* c, int32_t n, const uint8_t* const* src, int32_t stride) {
  if (!c || !src || n < 1 || n > c->opt.max_images || stride < c->opt.src_width) {
    set_error("svin_pre_upload: invalid arguments (num_images must be in [1, max_images], stride >= width)");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i < n; ++i)
    if (!src[i]) {
      set_error("svin_pre_upload: NULL image");
      return SVIN_ERR_INVALID_ARGUMENT;
    }
  SVIN_CUDA(cudaSetDevice(c->device));
  const svin::PreDev& p = c->p;
  c->pool->run(n, [&](int i) {
    for (int y = 0; y < p.sh; ++y)
      std::memcpy(c->h_src + ((size_t)i * p.sh + y) * p.spitch, src[i] + (size_t)y * stride, p.sw);
  });
  SVIN_CUDA(cudaEventRecord(c->ev[0], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->d_src, c->h_src, (size_t)n * p.sh * p.spitch, cudaMemcpyHostToDevice, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[1], c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->tm.h2d_ms = ms;
  c->tm.h2d_bytes = (int64_t)n * p.sh * p.spitch;
  c->n_images = n;
  return SVIN_OK;
}

static int pre_enqueue(svin_pre_ctx* c) {
  const svin::PreDev& p = c->p;
  const int n = c->n_images;
  const dim3 blk(128), grid((p.w + 127) / 128, p.h, n);
  cudaEvent_t* ev = c->ev;
  SVIN_CUDA(cudaEventRecord(ev[2], c->stream));
  const dim3 grid4((p.w + 4 * 128 - 1) / (4 * 128), p.h, n);
  if (p.mode == 1)
    svin::k_pre_decimate2<<<grid4, blk, 0, c->stream>>>(p);
  else if (p.mode == 2)
    svin::k_pre_bilinear<<<grid, blk, 0, c->stream>>>(p);
  else
    svin::k_pre_copy<<<grid, blk, 0, c->stream>>>(p);
  SVIN_CUDA(cudaEventRecord(ev[3], c->stream));
  const uint8_t* cur = p.a;
  uint8_t* other = p.b;
  c->tm.kernel_launches += 1;
  if (c->opt.median_filter) {
    svin::k_pre_median3<<<grid, blk, 0, c->stream>>>(p, cur, other);
    std::swap(const_cast<uint8_t*&>(cur), other);
    c->tm.kernel_launches += 1;
  }
  SVIN_CUDA(cudaEventRecord(ev[4], c->stream));
  if (p.method != SVIN_HIST_NONE) {
    const int T = p.tiles, th = p.method == SVIN_HIST_CLAHE ? p.th : p.h;
    SVIN_CUDA(cudaMemsetAsync(c->d_hist, 0, sizeof(unsigned) * (size_t)n * T * T * 256, c->stream));
    svin::k_pre_hist<<<dim3(T * T, (th + svin::kHistRows - 1) / svin::kHistRows, n), 256, 0, c->stream>>>(p, cur);
    SVIN_CUDA(cudaEventRecord(ev[5], c->stream));
    svin::k_pre_lut<<<dim3(T * T, n), 256, 0, c->stream>>>(p);
    SVIN_CUDA(cudaEventRecord(ev[6], c->stream));
    svin::k_pre_apply<<<dim3((p.h + svin::kApplyRows - 1) / svin::kApplyRows, n), 256, (size_t)T * T * 256 + sizeof(float2) * p.w, c->stream>>>(
        p, cur, other);
    std::swap(const_cast<uint8_t*&>(cur), other);
    c->tm.kernel_launches += 3;
  } else {
    SVIN_CUDA(cudaEventRecord(ev[5], c->stream));
    SVIN_CUDA(cudaEventRecord(ev[6], c->stream));
  }
  SVIN_CUDA(cudaEventRecord(ev[7], c->stream));
  SVIN_CUDA(cudaGetLastError());
  c->d_out = cur;
  return SVIN_OK;
}

int svin_pre_run(svin_pre_ctx* c) {
  if (!c || c->n_images < 1) {
    set_error("svin_pre_run: nothing uploaded");
    return SVIN_ERR_STATE;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const int rc = pre_enqueue(c);
  if (rc != SVIN_OK) return rc;
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[7]);
  c->tm.run_ms = ms;
  for (int k = 0; k < 5; ++k) {
    cudaEventElapsedTime(&ms, c->ev[2 + k], c->ev[3 + k]);
    c->tm.kernel_ms[k] = ms;
  }
  return SVIN_OK;
}

int svin_pre_download(svin_pre_ctx* c, uint8_t* const* dst) {
  if (!c || !dst || !c->d_out) {
    set_error("svin_pre_download: nothing to download");
    return SVIN_ERR_STATE;
  }
  SVIN_CUDA(cudaSetDevice(c->device));
  const svin::PreDev& p = c->p;
  const size_t per = (size_t)p.w * p.h;
  SVIN_CUDA(cudaEventRecord(c->ev[0], c->stream));
  SVIN_CUDA(cudaMemcpyAsync(c->h_dst, c->d_out, per * c->n_images, cudaMemcpyDeviceToHost, c->stream));
  SVIN_CUDA(cudaEventRecord(c->ev[1], c->stream));
  SVIN_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->tm.d2h_ms = ms;
  c->tm.d2h_bytes = (int64_t)(per * c->n_images);
  c->pool->run(c->n_images, [&](int i) {
    if (dst[i]) std::memcpy(dst[i], c->h_dst + per * i, per);
  });
  return SVIN_OK;
}

int svin_pre_process(svin_pre_ctx* c, int32_t n, const uint8_t* const* src, int32_t stride, uint8_t* const* dst) {
  int rc = svin_pre_upload(c, n, src, stride);
  if (rc != SVIN_OK) return rc;
  if ((rc = svin_pre_run(c)) != SVIN_OK) return rc;
  return svin_pre_download(c, dst);
}

const uint8_t* svin_pre_device_output(svin_pre_ctx* c) { return c ? c->d_out : nullptr; }

int svin_pre_timings(svin_pre_ctx* c, SvinPreTimings* out) {
  if (!c || !out) {
    set_error("svin_pre_timings: invalid arguments");
    return SVIN_ERR_INVALID_ARGUMENT;
  }
  *out = c->tm;
  return SVIN_OK;
}

}  // extern "C"
