// ConvTranspose3d k2 s2 (unets.deconvBlock, lib/network_factory/unets.py:42-58) as three small GEMMs on the warp-level
// tensor-core path (mma.sync m16n8k8 tf32, fp32 accumulate), 3xTF32 split so that products keep fp32 accuracy
// (hi = x & 0xffffe000, lo = (x - hi) & 0xffffe000; lo*hi + hi*lo + hi*hi).  The op has no halo and is HBM-bound once
// the arithmetic leaves the FP32 pipe (8x more output than input, 14-28 FLOP/B): no shared-memory staging of the
// activations at all -- every lane loads the fragment elements it owns straight from the planar tensors (32-byte sector
// granularity, every sector fully used) and splits them in registers; only the weights sit in shared memory.
//   forward:   Y[v][(co,pos)]  = sum_ci X[v][ci] W[ci][(co,pos)]      M = voxels, N = 8 Cout, K = Cin
//   data grad: dX[v][ci]       = sum_(co,pos) dY[v][(co,pos)] W[ci][(co,pos)]   M = voxels, N = Cin, K = 8 Cout
//   weight grad: dW[ci][(co,pos)] = sum_v X[ci][v] dY[v][(co,pos)]    M = Cin, N = 8 Cout, K = voxels
// with v a flattened input voxel, pos = (a, b, c) the position inside the 2x2x2 output cell, (co,pos) = co*8 + pos the
// weight tensor's own inner layout.  Included by deconv.cu.
//
// mma.sync.m16n8k8 tf32 fragments (g = lane >> 2, t = lane & 3):  A[16x8]: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// B[8x8]: b0 (k = t, n = g) b1 (k = t+4, n = g);  D[16x8]: d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1).

__device__ __forceinline__ void dm_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void dm_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
// d += A * B with fp32-grade products: small terms first
template <int NA>
__device__ __forceinline__ void dm_split_n(const float (&x)[NA], uint32_t (&hi)[NA], uint32_t (&lo)[NA]) {
#pragma unroll
  for (int i = 0; i < NA; ++i) dm_split(x[i], hi[i], lo[i]);
}
__device__ __forceinline__ void dm_mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2],
                                        const uint32_t (&bl)[2]) {
  dm_mma(d, al, bh);
  dm_mma(d, ah, bl);
  dm_mma(d, ah, bh);
}

constexpr int DM_FWD_THREADS = 128, DM_DGRAD_THREADS = 256;

// ---- forward: warp = 32 voxels (two M tiles) x the block's 256 output columns, in four chunks of 64 columns ----------
// grid (persistent, 8 Cout / 256, N); shared memory: W[CIN][256 + 8] (pitch 264: the four k rows of a B fragment start 8
// banks apart)
template <int CIN>
__global__ void __launch_bounds__(DM_FWD_THREADS) deconv_k2s2_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                                         int Cout, int D, int H, int W) {
  extern __shared__ __align__(16) float dm_smem[];
  constexpr int P = 256 + 8, KS = CIN / 8;
  const int N8 = 8 * Cout, nb0 = blockIdx.y * 256, n = blockIdx.z;
  for (int i = threadIdx.x; i < CIN * 256; i += blockDim.x) {
    const int k = i >> 8, c = i & 255;
    dm_smem[k * P + c] = w[(int64_t)k * N8 + nb0 + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t V = (int64_t)D * H * W, Vo = 8 * V;
  const int Ho = 2 * H, Wo = 2 * W;
  const float* xn = x + (int64_t)n * CIN * V;
  float* on = out + (int64_t)n * Cout * Vo;
  const int64_t ntiles = da_cdiv(V, 32);
  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; tile < ntiles; tile += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    // rows of this lane: v0 + {g, g+8, 16+g, 24+g}
    int64_t vr[4], obase[4];
    bool ok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      vr[r] = tile * 32 + r * 8 + g;
      ok[r] = vr[r] < V;
      const int64_t vv = ok[r] ? vr[r] : 0;
      const int xx = (int)(vv % W), yy = (int)((vv / W) % H), zz = (int)(vv / ((int64_t)W * H));
      obase[r] = ((int64_t)(2 * zz) * Ho + 2 * yy) * Wo + 2 * xx;
    }
    float a[KS][2][4];   // [k step][M tile][a0..a3]
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* p0 = xn + (int64_t)(ks * 8 + t) * V;
        const float* p1 = p0 + 4 * V;
        a[ks][mt][0] = ok[2 * mt] ? __ldg(p0 + vr[2 * mt]) : 0.f;
        a[ks][mt][1] = ok[2 * mt + 1] ? __ldg(p0 + vr[2 * mt + 1]) : 0.f;
        a[ks][mt][2] = ok[2 * mt] ? __ldg(p1 + vr[2 * mt]) : 0.f;
        a[ks][mt][3] = ok[2 * mt + 1] ? __ldg(p1 + vr[2 * mt + 1]) : 0.f;
      }
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {   // 64 output columns = 8 output channels
      float acc[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[2][4], al[2][4];
        dm_split_n<4>(a[ks][0], ah[0], al[0]);
        dm_split_n<4>(a[ks][1], ah[1], al[1]);
        const float* bp = dm_smem + (ks * 8 + t) * P + ch * 64 + g;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          uint32_t bh[2], bl[2];
          dm_split(bp[nt * 8], bh[0], bl[0]);
          dm_split(bp[nt * 8 + 4 * P], bh[1], bl[1]);
          dm_mma3(acc[0][nt], ah[0], al[0], bh, bl);
          dm_mma3(acc[1][nt], ah[1], al[1], bh, bl);
        }
      }
      // columns 2t, 2t+1 of N tile nt: output channel co, cell position (a, b, c = 0 / 1): one float2 per row
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int co = (nb0 >> 3) + ch * 8 + nt;
        const float bv = bias ? __ldg(bias + co) : 0.f;
        const int64_t off = (int64_t)co * Vo + (int64_t)(t >> 1) * Ho * Wo + (int64_t)(t & 1) * Wo;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (ok[2 * mt]) *reinterpret_cast<float2*>(on + off + obase[2 * mt]) = make_float2(acc[mt][nt][0] + bv, acc[mt][nt][1] + bv);
          if (ok[2 * mt + 1])
            *reinterpret_cast<float2*>(on + off + obase[2 * mt + 1]) = make_float2(acc[mt][nt][2] + bv, acc[mt][nt][3] + bv);
        }
      }
    }
  }
}

// ---- data gradient: warp = 16 * MT voxels x all CIN columns, K = 8 Cout walked one output channel (8 cell positions)
// per step with the next step's dY values already in flight ---------------------------------------------------------
// grid (persistent, 1, N); shared memory: W[CIN][8 Cout + 4] (pitch: the eight n rows of a B fragment start 4 banks apart)
template <int CIN, int MT>
__global__ void __launch_bounds__(DM_DGRAD_THREADS) deconv_k2s2_dgrad_mma_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                           float* __restrict__ dx, int Cout, int D, int H, int W) {
  extern __shared__ __align__(16) float dm_smem[];
  constexpr int NT = CIN / 8;
  const int N8 = 8 * Cout, P = N8 + 4, n = blockIdx.z;
  for (int i = threadIdx.x; i < CIN * N8; i += blockDim.x) {
    const int ci = i / N8, k = i - ci * N8;
    dm_smem[ci * P + k] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t V = (int64_t)D * H * W, Vo = 8 * V;
  const int Ho = 2 * H, Wo = 2 * W;
  const float* gn = dy + (int64_t)n * Cout * Vo;
  float* dn = dx + (int64_t)n * CIN * V;
  // A fragment columns of this lane: cell positions t (a = 0) and t + 4 (a = 1), b = t >> 1, c = t & 1
  const int64_t poff0 = (int64_t)(t >> 1) * Wo + (t & 1), poff1 = poff0 + (int64_t)Ho * Wo;
  const int64_t ntiles = da_cdiv(V, 16 * MT);
  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; tile < ntiles; tile += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    int64_t vr[2 * MT], obase[2 * MT];
    bool ok[2 * MT];
#pragma unroll
    for (int r = 0; r < 2 * MT; ++r) {
      vr[r] = tile * (16 * MT) + r * 8 + g;
      ok[r] = vr[r] < V;
      const int64_t vv = ok[r] ? vr[r] : 0;
      const int xx = (int)(vv % W), yy = (int)((vv / W) % H), zz = (int)(vv / ((int64_t)W * H));
      obase[r] = ((int64_t)(2 * zz) * Ho + 2 * yy) * Wo + 2 * xx;
    }
    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    float a[MT][4], an[MT][4];
    auto load_a = [&](int co, float (&q)[MT][4]) {
      const float* pc = gn + (int64_t)co * Vo;
      const bool live = co < Cout;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        q[mt][0] = (live && ok[2 * mt]) ? __ldg(pc + obase[2 * mt] + poff0) : 0.f;
        q[mt][1] = (live && ok[2 * mt + 1]) ? __ldg(pc + obase[2 * mt + 1] + poff0) : 0.f;
        q[mt][2] = (live && ok[2 * mt]) ? __ldg(pc + obase[2 * mt] + poff1) : 0.f;
        q[mt][3] = (live && ok[2 * mt + 1]) ? __ldg(pc + obase[2 * mt + 1] + poff1) : 0.f;
      }
    };
    load_a(0, a);
#pragma unroll 2
    for (int co = 0; co < Cout; ++co) {
      load_a(co + 1, an);
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) dm_split_n<4>(a[mt], ah[mt], al[mt]);
      const float* bp = dm_smem + g * P + co * 8 + t;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t bh[2], bl[2];
        dm_split(bp[nt * 8 * P], bh[0], bl[0]);
        dm_split(bp[nt * 8 * P + 4], bh[1], bl[1]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dm_mma3(acc[mt][nt], ah[mt], al[mt], bh, bl);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[mt][i] = an[mt][i];
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float* p0 = dn + (int64_t)(nt * 8 + 2 * t) * V;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (ok[2 * mt]) { p0[vr[2 * mt]] = acc[mt][nt][0]; p0[V + vr[2 * mt]] = acc[mt][nt][1]; }
        if (ok[2 * mt + 1]) { p0[vr[2 * mt + 1]] = acc[mt][nt][2]; p0[V + vr[2 * mt + 1]] = acc[mt][nt][3]; }
      }
    }
  }
}

// ---- weight gradient: block = all CIN rows x 256 columns (32 output channels), warp = (32 input channels, 8 output
// channels); K = voxels of the block's region, eight consecutive voxels of one input row per step (W % 8 == 0), the next
// step's operands in flight.  Also the bias gradient (column sums of dY, by the warps of the first row group). --------
// grid (regions, 8 Cout / 256, 1), block 32 * (CIN / 32) * 4 threads.  partials [region][CIN][8 Cout];
// bias_partials [region][Cout] nullable.
template <int CIN>
__global__ void __launch_bounds__(32 * (CIN / 32) * 4) deconv_k2s2_wgrad_mma_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                                    float* __restrict__ partials,
                                                                                    float* __restrict__ bias_partials, int N, int Cout, int D,
                                                                                    int H, int W, int steps_per_region) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int mr = warp >> 2, nc = warp & 3;                 // row group (32 input channels), column chunk (8 output channels)
  const int co0 = blockIdx.y * 32 + nc * 8;
  const int64_t V = (int64_t)D * H * W, Vo = 8 * V;
  const int Ho = 2 * H, Wo = 2 * W, W8 = W / 8;
  const int N8 = 8 * Cout;
  const int64_t nsteps = (int64_t)N * D * H * W8;
  const int64_t s0 = (int64_t)blockIdx.x * steps_per_region, s1 = min(nsteps, s0 + steps_per_region);
  // B fragment column g = cell position (a, b, c) = (g >> 2, (g >> 1) & 1, g & 1)
  const int64_t poff = (int64_t)(g >> 2) * Ho * Wo + (int64_t)((g >> 1) & 1) * Wo + (g & 1);
  const bool do_bias = bias_partials != nullptr && mr == 0;
  float acc[2][8][4], bsum[8];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    bsum[nt] = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  }
  struct Ops { float a[2][4]; float b[8][2]; };
  auto load_step = [&](int64_t s, Ops& q) {
    const bool live = s < s1;
    const int64_t ss = live ? s : s0;
    const int x8 = (int)(ss % W8);
    int64_t r = ss / W8;
    const int yy = (int)(r % H); r /= H;
    const int zz = (int)(r % D);
    const int n = (int)(r / D);
    const int64_t v = ((int64_t)zz * H + yy) * W + x8 * 8 + t;             // this lane's k rows: v, v + 4
    const float* px = x + ((int64_t)n * CIN + mr * 32 + g) * V + v;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      q.a[mt][0] = live ? __ldg(px + (int64_t)(mt * 16) * V) : 0.f;
      q.a[mt][1] = live ? __ldg(px + (int64_t)(mt * 16 + 8) * V) : 0.f;
      q.a[mt][2] = live ? __ldg(px + (int64_t)(mt * 16) * V + 4) : 0.f;
      q.a[mt][3] = live ? __ldg(px + (int64_t)(mt * 16 + 8) * V + 4) : 0.f;
    }
    const float* pg = dy + ((int64_t)n * Cout + co0) * Vo + ((int64_t)(2 * zz) * Ho + 2 * yy) * Wo + 2 * (x8 * 8 + t) + poff;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      q.b[nt][0] = live ? __ldg(pg + (int64_t)nt * Vo) : 0.f;
      q.b[nt][1] = live ? __ldg(pg + (int64_t)nt * Vo + 8) : 0.f;
    }
  };
  Ops cur, nxt;
  load_step(s0, cur);
  for (int64_t s = s0; s < s1; ++s) {
    load_step(s + 1, nxt);
    uint32_t ah[2][4], al[2][4];
    dm_split_n<4>(cur.a[0], ah[0], al[0]);
    dm_split_n<4>(cur.a[1], ah[1], al[1]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      uint32_t bh[2], bl[2];
      dm_split_n<2>(cur.b[nt], bh, bl);
      dm_mma3(acc[0][nt], ah[0], al[0], bh, bl);
      dm_mma3(acc[1][nt], ah[1], al[1], bh, bl);
      if (do_bias) bsum[nt] += cur.b[nt][0] + cur.b[nt][1];
    }
    cur = nxt;
  }
  float* pr = partials + (int64_t)blockIdx.x * CIN * N8;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int ci = mr * 32 + mt * 16 + g, col = (co0 + nt) * 8 + 2 * t;
      *reinterpret_cast<float2*>(pr + (int64_t)ci * N8 + col) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<float2*>(pr + (int64_t)(ci + 8) * N8 + col) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
    }
  if (do_bias) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float sres = warp_sum(bsum[nt]);   // all eight cell positions and the lane's k rows
      if (lane == 0) bias_partials[(int64_t)blockIdx.x * Cout + co0 + nt] = sres;
    }
  }
}
