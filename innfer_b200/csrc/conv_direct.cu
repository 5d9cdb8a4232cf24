#include "conv_direct.cuh"

#include <vector>

namespace innfer {

namespace {

constexpr int kTY = 8, kTX = 32;  // output pixels per block (rows x cols), one thread each

struct DirectParams {
  const float* in;
  int in_CT, in_chunk0, cin_chunks;
  int B, H, W, up;      // source dims
  int Ho, Wo;
  int ps, ps_a, ps_b;   // pixel-shuffle factor (1 = none) and this launch's output sub-pixel
  const float* w;       // [ci_pad][9][N]
  const float* bias;    // [N]
  float* out;
  int out_CT, out_chunk0, out_nchunks;
  int lrelu;
  float slope;
  const float* res1;
  int res1_CT, res1_chunk0;
  float alpha1;
  const float* res2;
  int res2_CT, res2_chunk0;
  float alpha2;
  int dil;              // dilation (taps at multiples of dil), <= MAXD of the instantiation
  int act_after_res;    // LeakyReLU after the residual adds
  int gate;             // res1 multiplies sigmoid(conv) instead of being added
  float* raw;           // optional pre-activation copy (layout of out)
  int raw_CT, raw_chunk0;
};

template <int N, int MAXD>
__global__ void __launch_bounds__(kTY* kTX)
conv_direct_kernel(const __grid_constant__ DirectParams p) {
  __shared__ float s_in[8][kTY + 2 * MAXD][kTX + 2 * MAXD];
  __shared__ __align__(16) float s_w[8][9][N];
  const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
  const int ox0 = blockIdx.x * kTX, oy0 = blockIdx.y * kTY, b = blockIdx.z;
  const int ox = ox0 + tx, oy = oy0 + ty;
  float acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;
  const size_t splane = (size_t)p.H * p.W;
  for (int cc = 0; cc < p.cin_chunks; ++cc) {
    __syncthreads();
    // halo tile in (possibly upsampled) output coordinates; zero outside [0,Ho)x[0,Wo)
    const int d = p.dil, hw = kTX + 2 * d;
    for (int i = threadIdx.x; i < (kTY + 2 * d) * hw; i += kTY * kTX) {
      const int hx = i % hw, hy = i / hw;
      const int uy = oy0 + hy - d, ux = ox0 + hx - d;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
      if (uy >= 0 && uy < p.Ho && ux >= 0 && ux < p.Wo) {
        const int sy = uy / p.up, sx = ux / p.up;
        const float4* src = reinterpret_cast<const float4*>(
            p.in + (((size_t)b * p.in_CT + p.in_chunk0 + cc) * splane + (size_t)sy * p.W + sx) * 8);
        a = src[0];
        c = src[1];
      }
      s_in[0][hy][hx] = a.x; s_in[1][hy][hx] = a.y; s_in[2][hy][hx] = a.z; s_in[3][hy][hx] = a.w;
      s_in[4][hy][hx] = c.x; s_in[5][hy][hx] = c.y; s_in[6][hy][hx] = c.z; s_in[7][hy][hx] = c.w;
    }
    for (int i = threadIdx.x; i < 8 * 9 * N; i += kTY * kTX)
      (&s_w[0][0][0])[i] = p.w[(size_t)cc * 8 * 9 * N + i];
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < 8; ++ci) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = s_in[ci][ty + (t / 3) * d][tx + (t % 3) * d];
        const float4* wv = reinterpret_cast<const float4*>(&s_w[ci][t][0]);
#pragma unroll
        for (int n4 = 0; n4 < N / 4; ++n4) {
          const float4 w4 = wv[n4];
          acc[4 * n4 + 0] = fmaf(v, w4.x, acc[4 * n4 + 0]);
          acc[4 * n4 + 1] = fmaf(v, w4.y, acc[4 * n4 + 1]);
          acc[4 * n4 + 2] = fmaf(v, w4.z, acc[4 * n4 + 2]);
          acc[4 * n4 + 3] = fmaf(v, w4.w, acc[4 * n4 + 3]);
        }
      }
    }
  }
  if (ox >= p.Wo || oy >= p.Ho) return;
  const size_t oplane = (size_t)p.Ho * p.Wo * p.ps * p.ps;
  const size_t opix = (size_t)(oy * p.ps + p.ps_a) * (p.Wo * p.ps) + (ox * p.ps + p.ps_b);
#pragma unroll
  for (int ch = 0; ch < N / 8; ++ch) {
    if (ch >= p.out_nchunks) break;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = acc[ch * 8 + e] + p.bias[ch * 8 + e];
      if (p.lrelu && !p.act_after_res) t = t > 0.f ? t : t * p.slope;
      f[e] = t;
    }
    if (p.gate == 2 && ch < N / 16) {   // merged PACnv: column c times sigmoid(column N/2 + c)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gv = acc[N / 2 + ch * 8 + e] + p.bias[N / 2 + ch * 8 + e];
        f[e] *= 1.f / (1.f + expf(-gv));
      }
    }
    if (p.res1) {
      const float4* r = reinterpret_cast<const float4*>(
          p.res1 + (((size_t)b * p.res1_CT + p.res1_chunk0 + ch) * oplane + opix) * 8);
      const float4 a = r[0], c = r[1];
      if (p.gate == 1) {   // res1 * sigmoid(conv)
        const float rv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = rv[e] * (1.f / (1.f + expf(-f[e])));
      } else {
        f[0] = f[0] * p.alpha1 + a.x; f[1] = f[1] * p.alpha1 + a.y; f[2] = f[2] * p.alpha1 + a.z; f[3] = f[3] * p.alpha1 + a.w;
        f[4] = f[4] * p.alpha1 + c.x; f[5] = f[5] * p.alpha1 + c.y; f[6] = f[6] * p.alpha1 + c.z; f[7] = f[7] * p.alpha1 + c.w;
      }
    }
    if (p.res2) {
      const float4* r = reinterpret_cast<const float4*>(
          p.res2 + (((size_t)b * p.res2_CT + p.res2_chunk0 + ch) * oplane + opix) * 8);
      const float4 a = r[0], c = r[1];
      f[0] = f[0] * p.alpha2 + a.x; f[1] = f[1] * p.alpha2 + a.y; f[2] = f[2] * p.alpha2 + a.z; f[3] = f[3] * p.alpha2 + a.w;
      f[4] = f[4] * p.alpha2 + c.x; f[5] = f[5] * p.alpha2 + c.y; f[6] = f[6] * p.alpha2 + c.z; f[7] = f[7] * p.alpha2 + c.w;
    }
    if (p.raw) {
      float4* o = reinterpret_cast<float4*>(p.raw + (((size_t)b * p.raw_CT + p.raw_chunk0 + ch) * oplane + opix) * 8);
      o[0] = make_float4(f[0], f[1], f[2], f[3]);
      o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    if (p.lrelu && p.act_after_res) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * p.slope;
    }
    float4* o = reinterpret_cast<float4*>(
        p.out + (((size_t)b * p.out_CT + p.out_chunk0 + ch) * oplane + opix) * 8);
    o[0] = make_float4(f[0], f[1], f[2], f[3]);
    o[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

}  // namespace

int conv_direct_upload(ConvLayer& L) {
  if (L.d_w32) return 0;
  const int N = L.N;
  const int nsets = L.pixel_shuffle ? L.nphase : 1;   // one filter set per output sub-pixel
  std::vector<float> t((size_t)nsets * L.Cin_pad * 9 * N, 0.f);
  for (int ph = 0; ph < nsets; ++ph)
    for (int co = 0; co < L.Cout; ++co) {
      const int row = L.pixel_shuffle ? co * L.nphase + ph : co;
      for (int ci = 0; ci < L.Cin; ++ci)
        for (int k = 0; k < 9; ++k)
          t[(((size_t)ph * L.Cin_pad + ci) * 9 + k) * N + co] = L.h_w32[((size_t)row * L.Cin + ci) * 9 + k];
    }
  cudaError_t e = cudaMalloc(&L.d_w32, t.size() * sizeof(float));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpy(L.d_w32, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice);
  return (int)e;
}

int conv_direct_run(const ConvLayer& L, ChunkView in, int B, int H, int W, ChunkView out,
                    int out_nchunks, const Epilogue& ep, cudaStream_t stream) {
  if (!L.d_w32) return -6;
  DirectParams p;
  p.in = reinterpret_cast<const float*>(in.base);
  p.in_CT = in.CT;
  p.in_chunk0 = in.chunk0;
  p.cin_chunks = L.Cin_pad / 8;
  p.B = B;
  p.H = H;
  p.W = W;
  p.up = L.pixel_shuffle ? 1 : L.up;
  p.Ho = H * p.up;
  p.Wo = W * p.up;
  p.ps = L.pixel_shuffle ? L.up : 1;
  p.ps_a = p.ps_b = 0;
  p.w = L.d_w32;
  p.bias = L.d_bias;
  p.out = reinterpret_cast<float*>(out.base);
  p.out_CT = out.CT;
  p.out_chunk0 = out.chunk0;
  p.out_nchunks = out_nchunks;
  p.lrelu = ep.lrelu ? 1 : 0;
  p.slope = ep.slope;
  p.res1 = reinterpret_cast<const float*>(ep.res1.base);
  p.res1_CT = ep.res1.CT;
  p.res1_chunk0 = ep.res1.chunk0;
  p.alpha1 = ep.alpha1;
  p.res2 = reinterpret_cast<const float*>(ep.res2.base);
  p.res2_CT = ep.res2.CT;
  p.res2_chunk0 = ep.res2.chunk0;
  p.alpha2 = ep.alpha2;
  p.dil = L.dil;
  p.act_after_res = ep.act_after_res ? 1 : 0;
  p.gate = ep.gate ? 1 : (ep.self_gate ? 2 : 0);
  if (ep.gate && !ep.res1.base) return -9;
  if (ep.res1_unact) return -9;   // fp16 paths only
  if (ep.self_gate && (ep.gate || out_nchunks > L.N / 16)) return -9;
  p.raw = reinterpret_cast<float*>(ep.raw_out.base);
  p.raw_CT = ep.raw_out.CT;
  p.raw_chunk0 = ep.raw_out.chunk0;
  if (L.dil != 1 && L.N != 32) return -8;   // dilated convs exist for 64 -> 32 only (PPON)
  dim3 grid((p.Wo + kTX - 1) / kTX, (p.Ho + kTY - 1) / kTY, B), block(kTY * kTX);
  const int nsets = L.pixel_shuffle ? L.nphase : 1;
  for (int ph = 0; ph < nsets; ++ph) {
    p.ps_a = ph / p.ps;
    p.ps_b = ph % p.ps;
    p.w = L.d_w32 + (size_t)ph * L.Cin_pad * 9 * L.N;
    p.bias = L.d_bias + (size_t)ph * L.N;
    switch (L.N) {
      case 16: conv_direct_kernel<16, 1><<<grid, block, 0, stream>>>(p); break;
      case 32:
        if (L.dil == 1) conv_direct_kernel<32, 1><<<grid, block, 0, stream>>>(p);
        else conv_direct_kernel<32, 8><<<grid, block, 0, stream>>>(p);
        break;
      case 64: conv_direct_kernel<64, 1><<<grid, block, 0, stream>>>(p); break;
      default: return -7;
    }
  }
  return (int)cudaGetLastError();
}

}  // namespace innfer
