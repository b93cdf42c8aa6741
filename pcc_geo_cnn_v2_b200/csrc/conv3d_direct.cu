// fp32 CUDA-core 3D convolution / transposed convolution, any cubic kernel size, stride 1 or 2,
// TensorFlow 'SAME' padding, fused bias + ReLU + residual add.
//
// Replaces Keras Conv3D / Conv3DTranspose (reference src/model_transforms.py:45-47,56-58,67-69,78-80,...)
// for the layers the tcgen05 path does not cover (C_in = 1 first layer, C_out = 1 last layer, the 9^3/5^3
// kernels of the V1 transforms) and serves as the fp32 on-device cross-check of the tensor-core kernels.
//
// One formulation covers every case: the output is split into `ncls` sub-pixel classes (1 for a forward
// conv or a stride-1 transposed conv, 8 parity classes for a stride-2 transposed conv).  Within a class
// every output voxel o (written at o*s_out + P) gathers from input voxel o*s_in + off_t for a per-dimension
// tap list (kernel index k_t, offset off_t).  No scatter, no atomics, deterministic.
//
// Thread mapping: VX output voxels (strided by blockDim so that loads/stores coalesce along W) x COB output
// channels per thread; taps outermost so the bounds logic is amortised over the C_in loop; weights are read
// through the read-only path as warp-uniform float4 broadcasts from the tap-major (k^3, Cin, Cout) layout.
#include "common.cuh"

namespace pccgeo {

struct DirectConvParams {
  const float* x;
  const float* w;
  const float* bias;
  const float* res;
  float* y;
  int N, Cin, Din, Hin, Win, Cout, Dout, Hout, Wout;
  int s_in, s_out, K, relu;
  int nt[2];
  int tk[2][9];
  int toff[2][9];
};

template <int COB, int VX>
__global__ void __launch_bounds__(128) conv3d_direct_kernel(const DirectConvParams p) {
  const int cls = blockIdx.z;
  const int Pz = (cls >> 2) & 1, Py = (cls >> 1) & 1, Px = cls & 1;
  const int co0 = blockIdx.y * COB;
  const int Dc = p.Dout / p.s_out, Hc = p.Hout / p.s_out, Wc = p.Wout / p.s_out;
  const long long total = (long long)p.N * Dc * Hc * Wc;
  const long long HWin = (long long)p.Hin * p.Win, DHWin = HWin * p.Din;

  int n[VX], oz[VX], oy[VX], ox[VX];
  bool valid[VX];
#pragma unroll
  for (int v = 0; v < VX; ++v) {
    long long L = ((long long)blockIdx.x * VX + v) * blockDim.x + threadIdx.x;
    valid[v] = L < total;
    if (!valid[v]) L = 0;
    ox[v] = (int)(L % Wc);
    L /= Wc;
    oy[v] = (int)(L % Hc);
    L /= Hc;
    oz[v] = (int)(L % Dc);
    n[v] = (int)(L / Dc);
  }
  float acc[VX][COB];
#pragma unroll
  for (int v = 0; v < VX; ++v)
#pragma unroll
    for (int c = 0; c < COB; ++c) acc[v][c] = 0.f;

  for (int jz = 0; jz < p.nt[Pz]; ++jz) {
    const int kz = p.tk[Pz][jz], dz = p.toff[Pz][jz];
    for (int jy = 0; jy < p.nt[Py]; ++jy) {
      const int ky = p.tk[Py][jy], dy = p.toff[Py][jy];
      for (int jx = 0; jx < p.nt[Px]; ++jx) {
        const int kx = p.tk[Px][jx], dx = p.toff[Px][jx];
        const float* wp = p.w + ((long long)((kz * p.K + ky) * p.K + kx) * p.Cin) * p.Cout + co0;
        long long off[VX];
        bool inb[VX];
        bool any = false;
#pragma unroll
        for (int v = 0; v < VX; ++v) {
          const int iz = oz[v] * p.s_in + dz, iy = oy[v] * p.s_in + dy, ix = ox[v] * p.s_in + dx;
          inb[v] = valid[v] && iz >= 0 && iz < p.Din && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
          off[v] = (long long)n[v] * p.Cin * DHWin + iz * HWin + (long long)iy * p.Win + ix;
          any |= inb[v];
        }
        if (!__any_sync(0xffffffffu, any)) continue;
        for (int ci = 0; ci < p.Cin; ++ci) {
          float xv[VX];
#pragma unroll
          for (int v = 0; v < VX; ++v) xv[v] = inb[v] ? __ldg(p.x + off[v] + ci * DHWin) : 0.f;
          float wv[COB];
          if (COB % 4 == 0) {
#pragma unroll
            for (int c = 0; c < COB; c += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(wp + (long long)ci * p.Cout + c));
              wv[c] = t.x; wv[c + 1] = t.y; wv[c + 2] = t.z; wv[c + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int c = 0; c < COB; ++c) wv[c] = __ldg(wp + (long long)ci * p.Cout + c);
          }
#pragma unroll
          for (int v = 0; v < VX; ++v)
#pragma unroll
            for (int c = 0; c < COB; ++c) acc[v][c] = fmaf(xv[v], wv[c], acc[v][c]);
        }
      }
    }
  }

  const long long HWout = (long long)p.Hout * p.Wout, DHWout = HWout * p.Dout;
#pragma unroll
  for (int v = 0; v < VX; ++v) {
    if (!valid[v]) continue;
    const int pz = oz[v] * p.s_out + Pz, py = oy[v] * p.s_out + Py, px = ox[v] * p.s_out + Px;
    const long long base = (long long)n[v] * p.Cout * DHWout + pz * HWout + (long long)py * p.Wout + px;
#pragma unroll
    for (int c = 0; c < COB; ++c) {
      float r = acc[v][c];
      if (p.bias) r += __ldg(p.bias + co0 + c);
      if (p.relu) r = fmaxf(r, 0.f);
      const long long idx = base + (long long)(co0 + c) * DHWout;
      if (p.res) r += __ldg(p.res + idx);
      p.y[idx] = r;
    }
  }
}

static int same_pad_before(int n, int k, int s) {
  int out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  return total / 2;
}

template <int COB, int VX>
static int launch(const DirectConvParams& p, int ncls, cudaStream_t st) {
  const int threads = 128;
  long long total = (long long)p.N * (p.Dout / p.s_out) * (p.Hout / p.s_out) * (p.Wout / p.s_out);
  long long bx = (total + (long long)threads * VX - 1) / ((long long)threads * VX);
  if (bx > 0x7fffffffLL) {
    set_error("conv3d_f32: grid too large");
    return PCCGEO_EINVAL;
  }
  dim3 grid((unsigned)bx, (unsigned)(p.Cout / COB), (unsigned)ncls);
  conv3d_direct_kernel<COB, VX><<<grid, threads, 0, st>>>(p);
  return check_launch("conv3d_direct_kernel");
}

}  // namespace pccgeo

extern "C" int pccgeo_conv3d_f32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                                 int n, int cin, int d, int h, int wd, int cout, int k, int stride, int transposed,
                                 int relu, void* stream) {
  using namespace pccgeo;
  PCCGEO_REQUIRE(x && w && y, "conv3d_f32: null pointer");
  PCCGEO_REQUIRE(n > 0 && cin > 0 && cout > 0 && d > 0 && h > 0 && wd > 0, "conv3d_f32: bad shape");
  PCCGEO_REQUIRE(k >= 1 && k <= 9 && (k & 1), "conv3d_f32: kernel size %d unsupported (odd, <= 9)", k);
  PCCGEO_REQUIRE(stride == 1 || stride == 2, "conv3d_f32: stride %d unsupported", stride);
  DirectConvParams p{};
  p.x = x; p.w = w; p.bias = bias; p.res = residual; p.y = y;
  p.N = n; p.Cin = cin; p.Din = d; p.Hin = h; p.Win = wd; p.Cout = cout; p.K = k; p.relu = relu;
  int ncls = 1;
  if (!transposed) {
    p.Dout = (d + stride - 1) / stride; p.Hout = (h + stride - 1) / stride; p.Wout = (wd + stride - 1) / stride;
    p.s_in = stride; p.s_out = 1;
    // SAME padding is computed per dimension; the tap lists are shared, so all dims must agree on pad_before.
    const int pb = same_pad_before(d, k, stride);
    PCCGEO_REQUIRE(same_pad_before(h, k, stride) == pb && same_pad_before(wd, k, stride) == pb,
                   "conv3d_f32: dims with different SAME padding are unsupported");
    // per-dim tap list: kernel index j reads input o*s + j - pb
    p.nt[0] = k; p.nt[1] = 0;
    for (int j = 0; j < k; ++j) { p.tk[0][j] = j; p.toff[0][j] = j - pb; }
  } else {
    p.Dout = d * stride; p.Hout = h * stride; p.Wout = wd * stride;
    p.s_in = 1; p.s_out = stride;
    ncls = stride == 1 ? 1 : 8;
    const int pb = same_pad_before(d * stride, k, stride);
    PCCGEO_REQUIRE(same_pad_before(h * stride, k, stride) == pb && same_pad_before(wd * stride, k, stride) == pb,
                   "conv3d_f32: dims with different SAME padding are unsupported");
    // output p = i*s + j - pb  =>  i = (p + pb - j)/s; for class parity P (p = o*s + P): off = (P + pb - j)/s
    for (int P = 0; P < stride; ++P) {
      int cnt = 0;
      for (int j = 0; j < k; ++j) {
        int num = P + pb - j;
        if (((num % stride) + stride) % stride != 0) continue;
        PCCGEO_REQUIRE(cnt < 9, "conv3d_f32: tap list overflow");
        p.tk[P][cnt] = j; p.toff[P][cnt] = num / stride;
        ++cnt;
      }
      p.nt[P] = cnt;
    }
    if (stride == 1) p.nt[1] = 0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (cout % 16 == 0) return launch<16, 4>(p, ncls, st);
  if (cout % 8 == 0) return launch<8, 4>(p, ncls, st);
  if (cout % 4 == 0) return launch<4, 4>(p, ncls, st);
  if (cout % 2 == 0) return launch<2, 4>(p, ncls, st);
  return launch<1, 4>(p, ncls, st);
}
