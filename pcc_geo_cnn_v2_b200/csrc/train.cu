// Backward / optimiser kernels of the tr_train.py path (reference src/model_types.py:327-369: focal loss + mbpov loss,
// two Adam optimisers; TF autodiff -> cuDNN backward kernels in the reference).  fp32, deterministic (fixed-order
// two-stage reductions, no fp atomics).  Data gradients of the convolutions reuse pccgeo_conv3d_f32 with the roles of
// conv / transposed conv swapped; this file adds what has no forward counterpart.
#include "common.cuh"

namespace pccgeo {

// ---- elementwise ---------------------------------------------------------------------------------------------------
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, float beta, float* __restrict__ out,
                             long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = alpha * a[i] + (b ? beta * b[i] : 0.f);
}

// ---- the same elementwise steps on the blocked bf16 hi/lo layout (term, N, C/8, voxels, 8): the tensor-core training path keeps the
// gradients in that layout between the layers of the backward pass (every conv kernel reads and writes it), so the ReLU mask, the
// residual adds and the bias gradients work on it directly instead of through a blocked -> fp32 -> blocked round trip per layer.
// `groups` = N * C/8 * voxels 16-byte items per term.
__device__ __forceinline__ uint32_t bf16x2_positive_mask(uint32_t w) {   // 0xffff per 16-bit lane that holds a value > 0
  const uint32_t lo = w & 0xffffu, hi = w >> 16;
  const uint32_t ml = ((lo & 0x8000u) == 0 && (lo & 0x7fffu) != 0) ? 0xffffu : 0u;
  const uint32_t mh = ((hi & 0x8000u) == 0 && (hi & 0x7fffu) != 0) ? 0xffff0000u : 0u;
  return ml | mh;
}
__global__ void relu_mask_blocked_kernel(const int4* __restrict__ g, const int4* __restrict__ y, int4* __restrict__ out, long long groups,
                                         int terms) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (long long)gridDim.x * blockDim.x) {
    const int4 yq = __ldg(y + i);   // the hi term of the activation decides: bf16_rn(v) > 0  <=>  v > 0
    const uint32_t m0 = bf16x2_positive_mask((uint32_t)yq.x), m1 = bf16x2_positive_mask((uint32_t)yq.y),
                   m2 = bf16x2_positive_mask((uint32_t)yq.z), m3 = bf16x2_positive_mask((uint32_t)yq.w);
    for (int t = 0; t < terms; ++t) {
      int4 q = __ldg(g + t * groups + i);
      q.x &= (int)m0; q.y &= (int)m1; q.z &= (int)m2; q.w &= (int)m3;
      out[t * groups + i] = q;
    }
  }
}
__device__ __forceinline__ void bf16x8_accumulate(const int4& q, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] += f.x;
    v[2 * k + 1] += f.y;
  }
}
__global__ void add_blocked_kernel(const int4* __restrict__ a, const int4* __restrict__ b, int4* __restrict__ out, long long groups, int terms) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (long long)gridDim.x * blockDim.x) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < terms; ++t) {
      bf16x8_accumulate(__ldg(a + t * groups + i), v);
      bf16x8_accumulate(__ldg(b + t * groups + i), v);
    }
    int4 qh, ql;
    uint32_t* qh32 = reinterpret_cast<uint32_t*>(&qh);
    uint32_t* ql32 = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
      __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
      qh32[k] = *reinterpret_cast<uint32_t*>(&hh);
      __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * k] - __bfloat162float(h0), v[2 * k + 1] - __bfloat162float(h1));
      ql32[k] = *reinterpret_cast<uint32_t*>(&ll);
    }
    out[i] = qh;
    if (terms == 2) out[groups + i] = ql;
  }
}
// db[c] = sum over blocks and voxels of (hi + lo); grid (chunks, C/8) -> partials[c * chunks + chunk] (bias_grad_finish_kernel adds them)
__global__ void bias_grad_blocked_kernel(const int4* __restrict__ g, double* __restrict__ partials, int N, int CG, long long V, int terms) {
  __shared__ double sm[32];
  const int cg = blockIdx.y;
  const long long groups = (long long)N * CG * V;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n = 0; n < N; ++n) {
    const int4* row = g + ((long long)n * CG + cg) * V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x)
      for (int t = 0; t < terms; ++t) bf16x8_accumulate(__ldg(row + t * groups + i), acc);
  }
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const double r = block_sum((double)acc[k], sm);
    if (threadIdx.x == 0) partials[(long long)(cg * 8 + k) * gridDim.x + blockIdx.x] = r;
  }
}

// d focal_loss / d y_pred (src/utils/focal_loss.py:5-12); K.clip passes gradient on [1e-3, .999] only; tf.where routes it
// through the selected branch only.  scale multiplies the result (lambda).
__global__ void focal_bwd_kernel(const float* __restrict__ xt, const float* __restrict__ xp, float gamma, float alpha, float scale,
                                 float* __restrict__ dxp, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float t = xt[i], p = xp[i];
    float g = 0.f;
    if (p >= 1e-3f && p <= .999f) {
      if (t == 1.f) {
        // -alpha * (1-p)^gamma * ln p
        g = -alpha * (-gamma * powf(1.f - p, gamma - 1.f) * logf(p) + powf(1.f - p, gamma) / p);
      } else if (t == 0.f) {
        // -(1-alpha) * p^gamma * ln(1-p)
        g = -(1.f - alpha) * (gamma * powf(p, gamma - 1.f) * logf(1.f - p) - powf(p, gamma) / (1.f - p));
      }
    }
    dxp[i] = scale * g;
  }
}

// ---- Gaussian conditional: d(c * sum ln p)/dv and /dsigma -------------------------------------------------------------
// p = Phi(u) - Phi(l), u = (.5-|v|)/s, l = (-.5-|v|)/s, s = lower_bound(sigma, smin), p = lower_bound(p, 1e-9).
// tfc lower_bound gradient ("identity_if_towards"): pass if input >= bound or the gradient is negative.
__global__ void gc_bwd_kernel(const float* __restrict__ v, const float* __restrict__ sigma, float smin, float c,
                              float* __restrict__ dv, float* __restrict__ dsigma, long long n) {
  const double kInvSqrt2 = 0.70710678118654752440, kInvSqrt2Pi = 0.39894228040143267794;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double sg = (double)sigma[i];
    const double s = sg > (double)smin ? sg : (double)smin;
    const double val = (double)v[i], a = fabs(val);
    const double u = (0.5 - a) / s, l = (-0.5 - a) / s;
    const double p = 0.5 * (erfc(-kInvSqrt2 * u) - erfc(-kInvSqrt2 * l));
    const double pb = p > 1e-9 ? p : 1e-9;
    double gp = (double)c / pb;               // dL/dp_bounded
    if (!(p >= 1e-9 || gp < 0.0)) gp = 0.0;   // lower_bound on the likelihood
    const double phu = kInvSqrt2Pi * exp(-0.5 * u * u), phl = kInvSqrt2Pi * exp(-0.5 * l * l);
    const double sgn = val > 0.0 ? 1.0 : (val < 0.0 ? -1.0 : 0.0);
    dv[i] = (float)(gp * (-(sgn) / s) * (phu - phl));
    double gs = gp * (-(u * phu - l * phl) / s);  // dL/ds
    if (!(sg >= (double)smin || gs < 0.0)) gs = 0.0;  // lower_bound on the scale
    dsigma[i] = (float)gs;
  }
}

// ---- entropy bottleneck: d(c * sum ln p)/dv and per-channel parameter gradients ---------------------------------------
// Parameters are the packed block of include/pccgeo.h (softplus'ed matrices M, biases B, tanh'ed factors F).  The kernel
// returns gradients w.r.t. M, B, F (43 values per channel; the host applies softplus'/tanh' for the raw variables).
struct EbFwd {
  float t[3][3];   // pre-gate activations of the three hidden layers
  float h[3][3];   // gated outputs  h = t + F * tanh(t)
  float out;
};

__device__ __forceinline__ float eb_forward(const float* __restrict__ q, float v, EbFwd& f) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    f.t[0][i] = q[i] * v + q[24 + i];
    f.h[0][i] = f.t[0][i] + q[34 + i] * tanhf(f.t[0][i]);
  }
#pragma unroll
  for (int l = 1; l < 3; ++l) {
    const float* M = q + (l == 1 ? 3 : 12);
    const float* B = q + (l == 1 ? 27 : 30);
    const float* F = q + (l == 1 ? 37 : 40);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      f.t[l][i] = M[i * 3] * f.h[l - 1][0] + M[i * 3 + 1] * f.h[l - 1][1] + M[i * 3 + 2] * f.h[l - 1][2] + B[i];
      f.h[l][i] = f.t[l][i] + F[i] * tanhf(f.t[l][i]);
    }
  }
  f.out = q[21] * f.h[2][0] + q[22] * f.h[2][1] + q[23] * f.h[2][2] + q[33];
  return f.out;
}

// accumulate g * d(out)/d(params) into acc[43] and return g * d(out)/dv
__device__ __forceinline__ float eb_backward(const float* __restrict__ q, float v, const EbFwd& f, float g, float* acc) {
  float dh[3], dt[3];
  // output layer: out = M3 . h2 + b3
#pragma unroll
  for (int i = 0; i < 3; ++i) { acc[21 + i] += g * f.h[2][i]; dh[i] = g * q[21 + i]; }
  acc[33] += g;
#pragma unroll
  for (int l = 2; l >= 1; --l) {
    const float* M = q + (l == 1 ? 3 : 12);
    const float* F = q + (l == 1 ? 37 : 40);
    const int mo = l == 1 ? 3 : 12, bo = l == 1 ? 27 : 30, fo = l == 1 ? 37 : 40;
    float dprev[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float th = tanhf(f.t[l][i]);
      acc[fo + i] += dh[i] * th;
      dt[i] = dh[i] * (1.f + F[i] * (1.f - th * th));
      acc[bo + i] += dt[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc[mo + i * 3 + j] += dt[i] * f.h[l - 1][j];
        dprev[j] += dt[i] * M[i * 3 + j];
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) dh[j] = dprev[j];
  }
  float dv = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float th = tanhf(f.t[0][i]);
    acc[34 + i] += dh[i] * th;
    dt[i] = dh[i] * (1.f + q[34 + i] * (1.f - th * th));
    acc[24 + i] += dt[i];
    acc[i] += dt[i] * v;
    dv += dt[i] * q[i];
  }
  return dv;
}

constexpr int kEbGradN = 44;  // gradients w.r.t. params [0:44) of the packed block ([43] = median: unused, kept 0)

// grid (chunks, C): block partial sums of the 44 gradients -> partials[(c * chunks + chunk) * 44 + k]
__global__ void eb_bwd_kernel(const float* __restrict__ vals, const float* __restrict__ params, float cscale, float* __restrict__ dv,
                              double* __restrict__ partials, int N, int C, int S) {
  __shared__ double sm[32];
  __shared__ float q[PCCGEO_EB_PARAM_STRIDE];
  const int c = blockIdx.y;
  if (threadIdx.x < PCCGEO_EB_PARAM_STRIDE) q[threadIdx.x] = params[c * PCCGEO_EB_PARAM_STRIDE + threadIdx.x];
  __syncthreads();
  float acc[kEbGradN];
#pragma unroll
  for (int k = 0; k < kEbGradN; ++k) acc[k] = 0.f;
  const long long total = (long long)N * S;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e / S), i = (int)(e % S);
    const long long idx = ((long long)n * C + c) * S + i;
    const float v = vals[idx];
    EbFwd fl, fu;
    const float lo = eb_forward(q, v - 0.5f, fl), up = eb_forward(q, v + 0.5f, fu);
    const float tt = lo + up;
    const float s = tt > 0.f ? -1.f : (tt < 0.f ? 1.f : 0.f);
    const float su = 1.f / (1.f + expf(-s * up)), sl = 1.f / (1.f + expf(-s * lo));
    const float diff = su - sl;
    const float p = fabsf(diff);
    const float pb = fmaxf(p, 1e-9f);
    float gp = cscale / pb;
    if (!(p >= 1e-9f || gp < 0.f)) gp = 0.f;
    const float sd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    const float gu = gp * sd * s * su * (1.f - su);     // dL/d up
    const float gl = -gp * sd * s * sl * (1.f - sl);    // dL/d lo
    float g = eb_backward(q, v + 0.5f, fu, gu, acc);
    g += eb_backward(q, v - 0.5f, fl, gl, acc);
    dv[idx] = g;
  }
#pragma unroll 1
  for (int k = 0; k < kEbGradN; ++k) {
    const double r = block_sum((double)acc[k], sm);
    if (threadIdx.x == 0) partials[((long long)c * gridDim.x + blockIdx.x) * kEbGradN + k] = r;
  }
}

__global__ void eb_bwd_finish_kernel(const double* __restrict__ partials, int chunks, float* __restrict__ dparams) {
  const int c = blockIdx.x, k = threadIdx.x;
  if (k >= kEbGradN) return;
  double s = 0.0;
  for (int j = 0; j < chunks; ++j) s += partials[((long long)c * chunks + j) * kEbGradN + k];
  dparams[c * kEbGradN + k] = (float)s;
}

// ---- convolution weight / bias gradients -----------------------------------------------------------------------------
// Unified form (conv and transposed conv): dW[t,ci,co] = sum_{n,b} x[n,ci, b*sx + ox_t] * g[n,co, b*sy + oy_t] over base
// positions b of a (Bd,Bh,Bw) grid; out-of-range positions contribute zero.  grid (taps, splits); each block reduces its
// slice of (n, b) into a Cin x Cout partial; a second kernel adds the splits in order.
struct WgradParams {
  const float* x;
  const float* g;
  float* partial;  // (taps, splits, Cin, Cout)
  int N, Cin, Cout, K;
  int Xd, Xh, Xw, Gd, Gh, Gw, Bd, Bh, Bw;
  int sx, sy, pb, transposed;
  int splits;
};

constexpr int WG_VC = 32;      // voxels per smem chunk
constexpr int WG_THREADS = 256;

__global__ void __launch_bounds__(WG_THREADS) wgrad_kernel(const WgradParams p) {
  extern __shared__ float smem[];
  float* xs = smem;                      // [Cin][WG_VC]
  float* gs = smem + p.Cin * WG_VC;      // [Cout][WG_VC]
  const int t = blockIdx.x, split = blockIdx.y;
  const int kz = t / (p.K * p.K), ky = (t / p.K) % p.K, kx = t % p.K;
  const int oxz = p.transposed ? 0 : kz - p.pb, oxy = p.transposed ? 0 : ky - p.pb, oxx = p.transposed ? 0 : kx - p.pb;
  const int ogz = p.transposed ? kz - p.pb : 0, ogy = p.transposed ? ky - p.pb : 0, ogx = p.transposed ? kx - p.pb : 0;
  const long long nb = (long long)p.Bd * p.Bh * p.Bw, total = nb * p.N;
  const long long per = (total + p.splits - 1) / p.splits;
  const long long lo = per * split, hi = lo + per < total ? lo + per : total;
  const int pairs = p.Cin * p.Cout;
  constexpr int MAXP = 16;  // pairs per thread (Cin*Cout <= 4096)
  // gradient sums cancel heavily (|sum| << sum|.|): each 32-voxel chunk is accumulated in fp32, the running sum in double
  double acc[MAXP];
#pragma unroll
  for (int i = 0; i < MAXP; ++i) acc[i] = 0.0;
  const long long XHW = (long long)p.Xh * p.Xw, XDHW = XHW * p.Xd, GHW = (long long)p.Gh * p.Gw, GDHW = GHW * p.Gd;
  for (long long base = lo; base < hi; base += WG_VC) {
    __syncthreads();
    // stage WG_VC base positions: a thread owns one voxel column (decoded once) and walks the channel rows
    {
      const int vv = threadIdx.x & (WG_VC - 1);
      const long long L = base + vv;
      const bool live = L < hi;
      long long xoff = -1, goff = -1;  // element offsets of channel 0, or -1 when out of range
      if (live) {
        const int n = (int)(L / nb);
        const long long r = L % nb;
        const int bz = (int)(r / ((long long)p.Bh * p.Bw)), by = (int)((r / p.Bw) % p.Bh), bx = (int)(r % p.Bw);
        int z = bz * p.sx + oxz, y = by * p.sx + oxy, x = bx * p.sx + oxx;
        if (z >= 0 && z < p.Xd && y >= 0 && y < p.Xh && x >= 0 && x < p.Xw)
          xoff = (long long)n * p.Cin * XDHW + z * XHW + (long long)y * p.Xw + x;
        z = bz * p.sy + ogz; y = by * p.sy + ogy; x = bx * p.sy + ogx;
        if (z >= 0 && z < p.Gd && y >= 0 && y < p.Gh && x >= 0 && x < p.Gw)
          goff = (long long)n * p.Cout * GDHW + z * GHW + (long long)y * p.Gw + x;
      }
      for (int ch = threadIdx.x / WG_VC; ch < p.Cin + p.Cout; ch += WG_THREADS / WG_VC) {
        float val = 0.f;
        if (ch < p.Cin) { if (xoff >= 0) val = p.x[xoff + (long long)ch * XDHW]; }
        else if (goff >= 0) val = p.g[goff + (long long)(ch - p.Cin) * GDHW];
        smem[ch * WG_VC + vv] = val;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int pr = threadIdx.x + i * WG_THREADS;
      if (pr < pairs) {
        const float* xr = xs + (pr / p.Cout) * WG_VC;
        const float* gr = gs + (pr % p.Cout) * WG_VC;
        float a = 0.f;
#pragma unroll
        for (int vv = 0; vv < WG_VC; ++vv) a = fmaf(xr[vv], gr[vv], a);
        acc[i] += (double)a;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int pr = threadIdx.x + i * WG_THREADS;
    if (pr < pairs) p.partial[((long long)t * p.splits + split) * pairs + pr] = (float)acc[i];
  }
}

// Register-tiled variant (channel counts are padded up to multiples of 4 with zero rows): a thread owns a
// 4 x 4 (ci, co) tile of one voxel slice, so each staged value feeds 4 FMAs from registers (the kernel above reads two
// shared-memory operands per FMA); voxel-major staging [v][(Cin+Cout)/4 + 1] float4 keeps the float4 reads of a warp on
// few distinct addresses (broadcast) and the staging stores 4-way at worst.  Same reduction contract: fp32 inside a chunk
// slice, double running sums, slices and splits added in a fixed order (deterministic).
template <int WT_VC>   // voxels per staged chunk (256 for narrow layers: fewer barriers / index decodes per FMA)
__global__ void __launch_bounds__(WG_THREADS, WT_VC == 256 ? 3 : 2) wgrad_tiled_kernel(const WgradParams p) {
  extern __shared__ float4 smem4[];
  const int C4i = (p.Cin + 3) / 4, C4o = (p.Cout + 3) / 4, C4 = C4i + C4o, ROW = C4 + 1;
  const int TP = C4i * C4o, VS = WG_THREADS / TP;          // tiles, voxel slices (TP divides 256)
  const int t = blockIdx.x, split = blockIdx.y;
  const int kz = t / (p.K * p.K), ky = (t / p.K) % p.K, kx = t % p.K;
  const int oxz = p.transposed ? 0 : kz - p.pb, oxy = p.transposed ? 0 : ky - p.pb, oxx = p.transposed ? 0 : kx - p.pb;
  const int ogz = p.transposed ? kz - p.pb : 0, ogy = p.transposed ? ky - p.pb : 0, ogx = p.transposed ? kx - p.pb : 0;
  const long long nb = (long long)p.Bd * p.Bh * p.Bw, total = nb * p.N;
  const long long per = (total + p.splits - 1) / p.splits;
  const long long lo = per * split, hi = lo + per < total ? lo + per : total;
  const long long XHW = (long long)p.Xh * p.Xw, XDHW = XHW * p.Xd, GHW = (long long)p.Gh * p.Gw, GDHW = GHW * p.Gd;
  const int tp = threadIdx.x % TP, vs = threadIdx.x / TP;
  const int cit = tp / C4o, cot = tp % C4o;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;
  for (long long base = lo; base < hi; base += WT_VC) {
    __syncthreads();
    {
      const int vv = threadIdx.x & (WT_VC - 1);
      const long long L = base + vv;
      long long xoff = -1, goff = -1;
      if (L < hi) {
        // total < 2^31 (host-checked): 32-bit index decode
        const uint32_t Lu = (uint32_t)L, nbu = (uint32_t)nb, hw = (uint32_t)(p.Bh * p.Bw);
        const int n = (int)(Lu / nbu);
        const uint32_t r = Lu - (uint32_t)n * nbu;
        const int bz = (int)(r / hw);
        const uint32_t r2 = r - (uint32_t)bz * hw;
        const int by = (int)(r2 / (uint32_t)p.Bw), bx = (int)(r2 - (uint32_t)by * (uint32_t)p.Bw);
        int z = bz * p.sx + oxz, y = by * p.sx + oxy, x = bx * p.sx + oxx;
        if (z >= 0 && z < p.Xd && y >= 0 && y < p.Xh && x >= 0 && x < p.Xw)
          xoff = (long long)n * p.Cin * XDHW + z * XHW + (long long)y * p.Xw + x;
        z = bz * p.sy + ogz; y = by * p.sy + ogy; x = bx * p.sy + ogx;
        if (z >= 0 && z < p.Gd && y >= 0 && y < p.Gh && x >= 0 && x < p.Gw)
          goff = (long long)n * p.Cout * GDHW + z * GHW + (long long)y * p.Gw + x;
      }
      for (int c4 = threadIdx.x / WT_VC; c4 < C4; c4 += WG_THREADS / WT_VC) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < C4i) {
          if (xoff >= 0) {
            const int c0 = 4 * c4;
            const float* q = p.x + xoff + (long long)c0 * XDHW;
            v.x = __ldg(q);
            if (c0 + 1 < p.Cin) v.y = __ldg(q + XDHW);
            if (c0 + 2 < p.Cin) v.z = __ldg(q + 2 * XDHW);
            if (c0 + 3 < p.Cin) v.w = __ldg(q + 3 * XDHW);
          }
        } else if (goff >= 0) {
          const int c0 = 4 * (c4 - C4i);
          const float* q = p.g + goff + (long long)c0 * GDHW;
          v.x = __ldg(q);
          if (c0 + 1 < p.Cout) v.y = __ldg(q + GDHW);
          if (c0 + 2 < p.Cout) v.z = __ldg(q + 2 * GDHW);
          if (c0 + 3 < p.Cout) v.w = __ldg(q + 3 * GDHW);
        }
        smem4[vv * ROW + c4] = v;
      }
    }
    __syncthreads();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    for (int v = vs; v < WT_VC; v += VS) {
      const float4 xv = smem4[v * ROW + cit], gv = smem4[v * ROW + C4i + cot];
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i * 4 + j] = fmaf(xa[i], ga[j], a[i * 4 + j]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] += (double)a[i];
  }
  // add the voxel slices in order: slice sums go through shared memory as doubles
  __syncthreads();
  double* red = reinterpret_cast<double*>(smem4);
  const int pairs = p.Cin * p.Cout;
  for (int s0 = 0; s0 < VS; ++s0) {
    if (vs == s0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cit * 4 + i >= p.Cin || cot * 4 + j >= p.Cout) continue;   // padding rows of the tile
          const int pr = (cit * 4 + i) * p.Cout + cot * 4 + j;
          red[pr] = (s0 == 0 ? 0.0 : red[pr]) + acc[i * 4 + j];
        }
    }
    __syncthreads();
  }
  for (int pr = threadIdx.x; pr < pairs; pr += WG_THREADS) p.partial[((long long)t * p.splits + split) * pairs + pr] = (float)red[pr];
}

__global__ void wgrad_finish_kernel(const float* __restrict__ partial, int splits, int pairs, float* __restrict__ dw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / pairs, pr = i % pairs;
    double s = 0.0;
    for (int j = 0; j < splits; ++j) s += (double)partial[(t * splits + j) * pairs + pr];
    dw[i] = (float)s;
  }
}

// One-channel ends of the V2 transforms (first layer 1 -> C, stride 2; last layer C -> 1, transposed): Cin*Cout = C pairs only, so
// the tap-per-block kernels above re-read the activations 27 times (10 ms for 16 -> 1 at 64^3 x 32).  Here a lane owns one base
// voxel of a row: the 27 neighbours of the one-channel tensor go to registers once, then every channel of the other tensor at
// that voxel (read exactly once, coalesced) feeds 27 FMAs.  A thread keeps 4 channels x 27 taps of fp32 sums over its ~100
// voxels; lanes are added by shuffles, warps / blocks by the finish kernel in double, in a fixed order.
constexpr int NARROW_BLOCKS = 592, NARROW_WARPS = 4, NARROW_CPT = 4;
struct NarrowParams {
  const float* multi;    // (N, C, Bd, Bh, Bw): indexed at the base voxel
  const float* single;   // (N, 1, Sd, Sh, Sw): indexed at base * ss + tap - pb
  float* partial;        // (blocks * warps, C, 27)
  int N, C, Bd, Bh, Bw, Sd, Sh, Sw, ss, pb;
};
__global__ void __launch_bounds__(NARROW_WARPS * 32) wgrad_narrow_kernel(const NarrowParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = blockIdx.y;
  const int xsegs = (p.Bw + 31) / 32;
  const long long rows = (long long)p.N * p.Bd * p.Bh * xsegs;
  float acc[NARROW_CPT][27];
#pragma unroll
  for (int c = 0; c < NARROW_CPT; ++c)
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[c][t] = 0.f;
  const long long BHW = (long long)p.Bh * p.Bw, BDHW = BHW * p.Bd, SHW = (long long)p.Sh * p.Sw, SDHW = SHW * p.Sd;
  for (long long row = (long long)blockIdx.x * NARROW_WARPS + warp; row < rows; row += (long long)gridDim.x * NARROW_WARPS) {
    const int xs = (int)(row % xsegs);
    long long r = row / xsegs;
    const int by = (int)(r % p.Bh); r /= p.Bh;
    const int bz = (int)(r % p.Bd);
    const int n = (int)(r / p.Bd);
    const int bx = xs * 32 + lane;
    const bool live = bx < p.Bw;
    float s[27];
    const float* sp = p.single + (long long)n * SDHW;
#pragma unroll
    for (int tz = 0; tz < 3; ++tz)
#pragma unroll
      for (int ty = 0; ty < 3; ++ty)
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
          const int z = bz * p.ss + tz - p.pb, y = by * p.ss + ty - p.pb, x = bx * p.ss + tx - p.pb;
          const bool ok = live && z >= 0 && z < p.Sd && y >= 0 && y < p.Sh && x >= 0 && x < p.Sw;
          s[(tz * 3 + ty) * 3 + tx] = ok ? __ldg(sp + z * SHW + (long long)y * p.Sw + x) : 0.f;
        }
    const float* mp = p.multi + ((long long)n * p.C + cg * NARROW_CPT) * BDHW + bz * BHW + (long long)by * p.Bw + bx;
#pragma unroll
    for (int c = 0; c < NARROW_CPT; ++c) {
      const float m = live ? __ldg(mp + c * BDHW) : 0.f;
#pragma unroll
      for (int t = 0; t < 27; ++t) acc[c][t] = fmaf(m, s[t], acc[c][t]);
    }
  }
  float* out = p.partial + ((long long)(blockIdx.x * NARROW_WARPS + warp) * p.C + cg * NARROW_CPT) * 27;
#pragma unroll
  for (int c = 0; c < NARROW_CPT; ++c)
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      float v = acc[c][t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) out[c * 27 + t] = v;
    }
}
// dw[t][c] = sum over the per-warp partials (warps, C, 27), in order, in double
__global__ void wgrad_narrow_finish_kernel(const float* __restrict__ partial, int nwarps, int C, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * 27) return;
  const int c = i / 27, t = i % 27;
  double s = 0.0;
  for (int j = 0; j < nwarps; ++j) s += (double)partial[((long long)j * C + c) * 27 + t];
  dw[t * C + c] = (float)s;
}

// db[c] = sum_{n,v} g[n,c,v]; grid (chunks, C) -> partials[c*chunks + chunk]; finish adds in order
__global__ void bias_grad_kernel(const float* __restrict__ g, double* __restrict__ partials, int N, int C, long long S) {
  __shared__ double sm[32];
  const int c = blockIdx.y;
  double acc = 0.0;
  if ((S & 3) == 0) {
    // 128-bit loads; fp32 inside one pass over a channel volume, double across passes
    const long long S4 = S >> 2;
    for (int n = 0; n < N; ++n) {
      const float4* row = reinterpret_cast<const float4*>(g + ((long long)n * C + c) * S);
      float a = 0.f;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(row + i);
        a += (v.x + v.y) + (v.z + v.w);
      }
      acc += (double)a;
    }
  } else {
    const long long total = (long long)N * S;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
      const long long n = e / S, i = e % S;
      acc += (double)g[(n * C + c) * S + i];
    }
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) partials[(long long)c * gridDim.x + blockIdx.x] = acc;
}
__global__ void bias_grad_finish_kernel(const double* __restrict__ partials, int chunks, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (int)gridDim.x * (int)blockDim.x) return;
  double s = 0.0;
  for (int j = 0; j < chunks; ++j) s += partials[(long long)c * chunks + j];
  db[c] = (float)s;
}

// TF1 AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v updates; theta -= lr_t * m / (sqrt(v) + eps)
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                            float lr_t, float b1, float b2, float eps, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i];
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

static int grid1d(long long n, int threads, int cap) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static int same_pb(int n, int k, int s) {
  int out = (n + s - 1) / s, total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  return total / 2;
}

}  // namespace pccgeo

using namespace pccgeo;

extern "C" int pccgeo_relu_bwd(const float* dy, const float* y, float* dx, long long n, void* stream) {
  PCCGEO_REQUIRE(dy && y && dx && n > 0, "relu_bwd: bad argument");
  relu_bwd_kernel<<<grid1d(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n);
  return check_launch("relu_bwd_kernel");
}

extern "C" int pccgeo_axpby(const float* a, const float* b, float alpha, float beta, float* out, long long n, void* stream) {
  PCCGEO_REQUIRE(a && out && n > 0, "axpby: bad argument");
  axpby_kernel<<<grid1d(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(a, b, alpha, beta, out, n);
  return check_launch("axpby_kernel");
}

extern "C" int pccgeo_relu_mask_blocked(const void* gb, const void* yb, void* out, long long groups, int terms, void* stream) {
  PCCGEO_REQUIRE(gb && yb && out && groups > 0 && (terms == 1 || terms == 2), "relu_mask_blocked: bad argument");
  relu_mask_blocked_kernel<<<grid1d(groups, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>((const int4*)gb, (const int4*)yb, (int4*)out, groups, terms);
  return check_launch("relu_mask_blocked_kernel");
}

extern "C" int pccgeo_add_blocked(const void* a, const void* b, void* out, long long groups, int terms, void* stream) {
  PCCGEO_REQUIRE(a && b && out && groups > 0 && (terms == 1 || terms == 2), "add_blocked: bad argument");
  add_blocked_kernel<<<grid1d(groups, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>((const int4*)a, (const int4*)b, (int4*)out, groups, terms);
  return check_launch("add_blocked_kernel");
}

constexpr int kBiasBlockedChunks = 148 * 8;
extern "C" size_t pccgeo_bias_grad_blocked_ws_doubles(int c) { return (size_t)((c + 15) / 16) * 16 * kBiasBlockedChunks; }

extern "C" int pccgeo_bias_grad_blocked(const void* gb, float* db, double* ws, int n, int c, long long spatial, int terms, void* stream) {
  PCCGEO_REQUIRE(gb && db && ws && n > 0 && c > 0 && spatial > 0 && (terms == 1 || terms == 2), "bias_grad_blocked: bad argument");
  const int cg = ((c + 15) / 16) * 2;   // the blocked layout pads the channels to a multiple of 16
  int chunks = kBiasBlockedChunks / cg;   // ~ 8 CTAs per SM in total
  const long long want = (spatial + 255) / 256;
  if (chunks > want) chunks = (int)want;
  if (chunks < 1) chunks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  bias_grad_blocked_kernel<<<dim3(chunks, cg), 256, 0, st>>>((const int4*)gb, ws, n, cg, spatial, terms);
  int rc = check_launch("bias_grad_blocked_kernel");
  if (rc) return rc;
  bias_grad_finish_kernel<<<1, c, 0, st>>>(ws, chunks, db);
  return check_launch("bias_grad_finish_kernel");
}

extern "C" int pccgeo_focal_loss_bwd(const float* x_true, const float* x_pred, float gamma, float alpha, float scale, float* dx_pred,
                                     long long n, void* stream) {
  PCCGEO_REQUIRE(x_true && x_pred && dx_pred && n > 0, "focal_loss_bwd: bad argument");
  focal_bwd_kernel<<<grid1d(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x_true, x_pred, gamma, alpha, scale, dx_pred, n);
  return check_launch("focal_bwd_kernel");
}

extern "C" int pccgeo_gc_likelihood_bwd(const float* values, const float* sigma, float scale_min, float c, float* dvalues,
                                        float* dsigma, long long n, void* stream) {
  PCCGEO_REQUIRE(values && sigma && dvalues && dsigma && n > 0, "gc_likelihood_bwd: bad argument");
  gc_bwd_kernel<<<grid1d(n, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(values, sigma, scale_min, c, dvalues, dsigma, n);
  return check_launch("gc_bwd_kernel");
}

extern "C" size_t pccgeo_eb_bwd_ws_doubles(int c) { return (size_t)c * 16 * kEbGradN; }

extern "C" int pccgeo_eb_likelihood_bwd(const float* values, const float* eb_params, float c, float* dvalues, float* dparams,
                                        double* ws, int n, int ch, int spatial, void* stream) {
  PCCGEO_REQUIRE(values && eb_params && dvalues && dparams && ws && n > 0 && ch > 0 && spatial > 0, "eb_likelihood_bwd: bad argument");
  const int chunks = 16;
  cudaStream_t st = (cudaStream_t)stream;
  eb_bwd_kernel<<<dim3(chunks, ch), 128, 0, st>>>(values, eb_params, c, dvalues, ws, n, ch, spatial);
  int rc = check_launch("eb_bwd_kernel");
  if (rc) return rc;
  eb_bwd_finish_kernel<<<ch, 64, 0, st>>>(ws, chunks, dparams);
  return check_launch("eb_bwd_finish_kernel");
}

extern "C" size_t pccgeo_wgrad_ws_floats(int cin, int cout, int k) {
  const int taps = k * k * k;
  int splits = (592 + taps - 1) / taps;
  size_t need = (size_t)taps * splits * cin * cout;
  if (k == 3 && (cin == 1 || cout == 1)) {
    const size_t narrow = (size_t)NARROW_BLOCKS * NARROW_WARPS * (cin * cout) * 27;
    if (narrow > need) need = narrow;
  }
  return need;
}

extern "C" int pccgeo_conv3d_wgrad_f32(const float* x, const float* g, float* dw, float* ws, int n, int cin, int d, int h, int wd,
                                       int cout, int k, int stride, int transposed, void* stream) {
  PCCGEO_REQUIRE(x && g && dw && ws && n > 0 && cin > 0 && cout > 0, "conv3d_wgrad: bad argument");
  PCCGEO_REQUIRE(cin * cout <= 4096, "conv3d_wgrad: Cin*Cout %d > 4096", cin * cout);
  PCCGEO_REQUIRE(k >= 1 && k <= 9 && (stride == 1 || stride == 2), "conv3d_wgrad: unsupported geometry");
  WgradParams p{};
  p.x = x; p.g = g; p.partial = ws; p.N = n; p.Cin = cin; p.Cout = cout; p.K = k; p.transposed = transposed;
  p.Xd = d; p.Xh = h; p.Xw = wd;
  if (!transposed) {
    p.Gd = (d + stride - 1) / stride; p.Gh = (h + stride - 1) / stride; p.Gw = (wd + stride - 1) / stride;
    p.Bd = p.Gd; p.Bh = p.Gh; p.Bw = p.Gw; p.sx = stride; p.sy = 1;
    p.pb = same_pb(d, k, stride);
    PCCGEO_REQUIRE(same_pb(h, k, stride) == p.pb && same_pb(wd, k, stride) == p.pb, "conv3d_wgrad: dims with different SAME padding");
  } else {
    p.Gd = d * stride; p.Gh = h * stride; p.Gw = wd * stride;
    p.Bd = d; p.Bh = h; p.Bw = wd; p.sx = 1; p.sy = stride;
    p.pb = same_pb(d * stride, k, stride);
  }
  const int taps = k * k * k;
  p.splits = (592 + taps - 1) / taps;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  {
    const int C = cin * cout;
    if (k == 3 && ((!transposed && cin == 1) || (transposed && cout == 1)) && C % NARROW_CPT == 0 && C >= NARROW_CPT) {
      NarrowParams q{};
      q.multi = transposed ? x : g; q.single = transposed ? g : x; q.partial = ws;
      q.N = n; q.C = C; q.Bd = p.Bd; q.Bh = p.Bh; q.Bw = p.Bw;
      q.Sd = transposed ? p.Gd : p.Xd; q.Sh = transposed ? p.Gh : p.Xh; q.Sw = transposed ? p.Gw : p.Xw;
      q.ss = stride; q.pb = p.pb;
      const long long rows = (long long)n * p.Bd * p.Bh * ((p.Bw + 31) / 32);
      int blocks = (int)((rows + NARROW_WARPS - 1) / NARROW_WARPS);
      if (blocks > NARROW_BLOCKS) blocks = NARROW_BLOCKS;
      wgrad_narrow_kernel<<<dim3(blocks, C / NARROW_CPT), NARROW_WARPS * 32, 0, st>>>(q);
      rc = check_launch("wgrad_narrow_kernel");
      if (rc) return rc;
      wgrad_narrow_finish_kernel<<<(C * 27 + 127) / 128, 128, 0, st>>>(ws, blocks * NARROW_WARPS, C, dw);
      return check_launch("wgrad_narrow_finish_kernel");
    }
  }
  const int tp = ((cin + 3) / 4) * ((cout + 3) / 4);
  if (tp <= WG_THREADS && WG_THREADS % tp == 0) {
    const int c4 = (cin + 3) / 4 + (cout + 3) / 4;
    const bool wide = c4 > 16;   // 32- and 64-channel layers: 128-voxel chunks; narrower ones: 256
    const int vc = wide ? 128 : 256;
    PCCGEO_REQUIRE((long long)p.N * p.Bd * p.Bh * p.Bw < (1LL << 31), "conv3d_wgrad: too many positions for the 32-bit index decode");
    size_t smem = (size_t)vc * (c4 + 1) * sizeof(float4);
    const size_t red = (size_t)cin * cout * sizeof(double);
    if (red > smem) smem = red;
    static bool attr = false;
    if (!attr) {
      PCCGEO_CUDA(cudaFuncSetAttribute(wgrad_tiled_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      PCCGEO_CUDA(cudaFuncSetAttribute(wgrad_tiled_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    if (wide) wgrad_tiled_kernel<128><<<dim3(taps, p.splits), WG_THREADS, smem, st>>>(p);
    else wgrad_tiled_kernel<256><<<dim3(taps, p.splits), WG_THREADS, smem, st>>>(p);
    rc = check_launch("wgrad_tiled_kernel");
  } else {
    const size_t smem = (size_t)(cin + cout) * WG_VC * sizeof(float);
    wgrad_kernel<<<dim3(taps, p.splits), WG_THREADS, smem, st>>>(p);
    rc = check_launch("wgrad_kernel");
  }
  if (rc) return rc;
  const long long total = (long long)taps * cin * cout;
  wgrad_finish_kernel<<<grid1d(total, 256, 148 * 8), 256, 0, st>>>(ws, p.splits, cin * cout, dw, total);
  return check_launch("wgrad_finish_kernel");
}

extern "C" int pccgeo_bias_grad_f32(const float* g, float* db, double* ws, int n, int c, long long spatial, void* stream) {
  PCCGEO_REQUIRE(g && db && ws && n > 0 && c > 0 && spatial > 0, "bias_grad: bad argument");
  int chunks = kReduceBlocks / c;
  PCCGEO_REQUIRE(chunks >= 1, "bias_grad: too many channels");
  {
    const long long want = (spatial / 4 + 255) / 256;   // blocks that still get a full 128-bit pass
    if (chunks > want) chunks = want < 1 ? 1 : (int)want;
  }
  cudaStream_t st = (cudaStream_t)stream;
  bias_grad_kernel<<<dim3(chunks, c), 256, 0, st>>>(g, ws, n, c, spatial);
  int rc = check_launch("bias_grad_kernel");
  if (rc) return rc;
  bias_grad_finish_kernel<<<1, c, 0, st>>>(ws, chunks, db);
  return check_launch("bias_grad_finish_kernel");
}

extern "C" int pccgeo_adam_step(float* theta, const float* grad, float* m, float* v, float lr, float beta1, float beta2, float eps,
                                long long step, long long n, void* stream) {
  PCCGEO_REQUIRE(theta && grad && m && v && n > 0 && step >= 1, "adam_step: bad argument");
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  adam_kernel<<<grid1d(n, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(theta, grad, m, v, (float)lr_t, beta1, beta2, eps, n);
  return check_launch("adam_kernel");
}
