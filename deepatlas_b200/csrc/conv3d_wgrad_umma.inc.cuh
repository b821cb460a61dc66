// Tensor-core (tcgen05, 3xTF32) weight gradient of the k3 s1 p1 convolutions.  Included by conv3d.cu.
//
//   dW[co][ci][kz][ky][kx] = sum_q dY[co][q] * X[ci][q + (kz-1, ky-1, kx-1)]
// as ONE GEMM per (16 ci x 16 co) block whose accumulator never leaves TMEM until the CTA has walked its whole region:
//   D[(t, ci)][(kz, ky, co)] += sum_k A[(t, ci)][k] * B[(kz, ky, co)][k],      k = (x, ye): 2 x-positions x 4 y-rows per MMA
//   A[(t, ci)][(x, ye)]      = X [ci][z][y0+ye][x0 - 1 + x + t]                 t = kx
//   B[(kz,ky,co)][(x, ye)]   = dY[co][z-kz+1][y0+ye-ky+1][x0 + x]
// * The kx shift costs nothing: the A tile is stored [x-position][ci 16][4 y-rows] (256 B per position), so MMA row
//   16*t + ci simply lands t positions further (SBO = 128 B, LBO = 256 B alias each other on purpose).  Rows t = 3..7 of
//   the 128-row MMA compute garbage that is never read.
// * ky and kz are folded into N = 144: the producers replicate each dY value into the nine (kz, ky) row groups of the
//   B tile (9 x 16 co rows per x-position); 1.2 KB of shared-memory stores per position against 30 MMA-cycles.
// * fp32 accuracy from three TF32 MMAs per K step (hi/lo split by the producers), as in conv3d_umma_kernel.
// * No epilogue per tile: 24 MMAs per 64-position tile, two-stage producer/MMA pipeline, one TMEM read at the end that
//   writes this CTA's partial sums (fixed-order region reduce afterwards, deterministic).
// Warp roles: 0 = MMA issue + TMEM allocation, 1..15 = producers (B tasks on threads 0..383, A tasks on 384..479 of the
// producer set), warps 4 and 5 read the accumulator at the end.

constexpr int WU_XT = 16, WU_YT = 4;             // tile: 16 x-positions x 4 y-rows of one plane
constexpr int WU_AC = WU_XT + 8;                 // A chunks per stage (x0-1 .. x0+XT+6)
constexpr int WU_N = 144;                        // (kz, ky, co)
constexpr int WU_A_BYTES = WU_AC * 256;          // per hi / lo half
constexpr int WU_BP = WU_N * 16 + 16;             // B pitch per x-position: +16 B so that lanes along x hit different banks
constexpr int WU_B_BYTES = WU_XT * WU_BP;        // 37120 per half
constexpr int WU_STAGE_BYTES = 2 * WU_A_BYTES + 2 * WU_B_BYTES;  // 86528
constexpr int WU_SMEM_BYTES = 2 * WU_STAGE_BYTES + 128;
constexpr int WU_THREADS = 512, WU_NPROD = 480;
constexpr int WU_BTASKS = WU_XT * 3 * 16, WU_ATASKS = WU_AC * 16;  // 768, 384
static_assert(WU_BTASKS == 2 * 384 && WU_ATASKS == 4 * 96, "producer mapping");

struct WgUmmaArgs {
  const float* x;   // [N][C][D][H][W]   halo side
  const float* dy;  // [N][Cout][D][H][W]
  float* partials;  // [region][Cout_total][Cin_total][27]
  int N, C, ci_off, Cin_total, Cout, co_off;
  int64_t region_stride;
  int D, H, W;
  int tiles_x, tiles_y, tiles_per_region, ntiles, nCoB;
};

__global__ void __launch_bounds__(WU_THREADS, 1) conv3d_wgrad_umma_kernel(WgUmmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2], empty[2], done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cob = blockIdx.x % a.nCoB, cib = blockIdx.x / a.nCoB;
  const int region = blockIdx.y;
  const int64_t HW = (int64_t)a.H * a.W, V = HW * a.D;
  const int t0 = region * a.tiles_per_region, t1 = min(a.ntiles, t0 + a.tiles_per_region);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], WU_NPROD); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // =============================== MMA issue ===============================
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(WU_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc0 = umma_desc(0, 256, 128), bdesc0 = umma_desc(0, WU_BP, 128);
    const uint32_t base_s = smem_u32(smem);
    uint32_t first = 1;
    for (int t = t0; t < t1; ++t) {
      const int k = t - t0, s = k & 1;
      mbar_wait(&full[s], (k >> 1) & 1);
      tc_fence_after();
      if (elected) {
        const uint32_t a_hi = base_s + (uint32_t)s * WU_STAGE_BYTES, a_lo = a_hi + WU_A_BYTES;
        const uint32_t b_hi = a_hi + 2 * WU_A_BYTES, b_lo = b_hi + WU_B_BYTES;
#pragma unroll
        for (int ks = 0; ks < WU_XT / 2; ++ks) {
          const uint32_t ao = (uint32_t)(2 * ks) * 256, bo = (uint32_t)(2 * ks) * WU_BP;
          umma_tf32(tmem, umma_desc_at(adesc0, a_hi + ao), umma_desc_at(bdesc0, b_hi + bo), idesc, first ? 0u : 1u);
          first = 0;
          umma_tf32(tmem, umma_desc_at(adesc0, a_lo + ao), umma_desc_at(bdesc0, b_hi + bo), idesc, 1u);
          umma_tf32(tmem, umma_desc_at(adesc0, a_hi + ao), umma_desc_at(bdesc0, b_lo + bo), idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elected) umma_commit(&done);
    __syncwarp();
  } else if (threadIdx.x < 32 + 384) {
    // =============================== B producers: (x, co, kz) fixed per (thread, j) ===============================
    const int tp = threadIdx.x - 32;
    const float* base[2];
    int tkz[2], tco[2];
    bool tok[2];
    const int tx = tp & 15;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tb = tp + 384 * j;
      tco[j] = (tb >> 4) & 15; tkz[j] = tb >> 8;
      tok[j] = cob * 16 + tco[j] < a.Cout;
      base[j] = a.dy + (int64_t)(cob * 16 + tco[j]) * V + tx;
    }
    float v[2][6], vn[2][6];
    auto gather = [&](int t, float (&o)[2][6]) {
      int tb = t;
      const int bx = tb % a.tiles_x; tb /= a.tiles_x;
      const int by = tb % a.tiles_y; tb /= a.tiles_y;
      const int z = tb % a.D;
      const int n = tb / a.D;
      const int x0 = bx * WU_XT, y0 = by * WU_YT;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int zz = z - tkz[j] + 1;
        const bool ok = tok[j] && zz >= 0 && zz < a.D && x0 + tx < a.W;
        const float* p = base[j] + (int64_t)n * a.Cout * V + (int64_t)zz * HW + x0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const int gy = y0 - 1 + r;
          o[j][r] = (ok && gy >= 0 && gy < a.H) ? __ldg(p + (int64_t)gy * a.W) : 0.f;
        }
      }
    };
    if (t0 < t1) gather(t0, vn);
    for (int t = t0; t < t1; ++t) {
      const int k = t - t0, s = k & 1, use = k >> 1;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int r = 0; r < 6; ++r) v[j][r] = vn[j][r];
      if (t + 1 < t1) gather(t + 1, vn);  // next tile's loads fly while this one is split and stored
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
      uint8_t* st = smem + s * WU_STAGE_BYTES + 2 * WU_A_BYTES;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {  // dY row y0 + e - ky + 1  =  v[e - ky + 2]
            const float val = v[j][e - ky + 2];
            h[e] = __uint_as_float(__float_as_uint(val) & 0xffffe000u);
            l[e] = val - h[e];
          }
          const int off = tx * WU_BP + ((tkz[j] * 3 + ky) * 16 + tco[j]) * 16;
          *reinterpret_cast<float4*>(st + off) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(st + WU_B_BYTES + off) = make_float4(l[0], l[1], l[2], l[3]);
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  } else {
    // =============================== A producers: (xc, ci) fixed per (thread, j) ===============================
    // lanes 0..7 of every quarter warp hold eight different ci (16 B apart in the tile): conflict-free 16-byte stores
    const int tp = threadIdx.x - 32 - 384, w = tp >> 5, l5 = tp & 31;
    const float* base[4];
    int txc[4], tci[4];
    bool tok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tci[j] = (l5 & 7) + 8 * (j & 1);
      txc[j] = (l5 >> 3) + 4 * (w + 3 * (j >> 1));
      tok[j] = cib * 16 + tci[j] < a.C && txc[j] < WU_XT + 2;  // chunks beyond feed only the unused rows t >= 3
      base[j] = a.x + (int64_t)(cib * 16 + tci[j]) * V + txc[j] - 1;
    }
    float v[4][4], vn[4][4];
    auto gather = [&](int t, float (&o)[4][4]) {
      int tb = t;
      const int bx = tb % a.tiles_x; tb /= a.tiles_x;
      const int by = tb % a.tiles_y; tb /= a.tiles_y;
      const int z = tb % a.D;
      const int n = tb / a.D;
      const int x0 = bx * WU_XT, y0 = by * WU_YT;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gx = x0 - 1 + txc[j];
        const bool ok = tok[j] && gx >= 0 && gx < a.W;
        const float* p = base[j] + (int64_t)n * a.C * V + (int64_t)z * HW + x0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int gy = y0 + r;
          o[j][r] = (ok && gy < a.H) ? __ldg(p + (int64_t)gy * a.W) : 0.f;
        }
      }
    };
    if (t0 < t1) gather(t0, vn);
    for (int t = t0; t < t1; ++t) {
      const int k = t - t0, s = k & 1, use = k >> 1;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) v[j][r] = vn[j][r];
      if (t + 1 < t1) gather(t + 1, vn);
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
      uint8_t* st = smem + s * WU_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h[e] = __uint_as_float(__float_as_uint(v[j][e]) & 0xffffe000u);
          l[e] = v[j][e] - h[e];
        }
        const int off = (txc[j] * 16 + tci[j]) * 16;
        *reinterpret_cast<float4*>(st + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(st + WU_A_BYTES + off) = make_float4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  }

  // ---- final read of the accumulator: rows 16*t + ci (t = kx < 3), columns (kz*3+ky)*16 + co ----
  if (warp == 4 || warp == 5) {
    mbar_wait(&done, 0);
    tc_fence_after();
    const int r = (warp - 4) * 32 + lane;  // TMEM lane = accumulator row; warp 4 -> lanes 0..31, warp 5 -> 32..63
    const int kx = r >> 4, ci = cib * 16 + (r & 15);
    float* pr = a.partials + (int64_t)region * a.region_stride;
#pragma unroll 1
    for (int g = 0; g < 9; ++g) {
      float v16[16];
      tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 16), v16);
      tmem_ld_wait();
      if (kx < 3 && ci < a.C && t0 < t1) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int co = cob * 16 + c;
          if (co < a.Cout) pr[((int64_t)(a.co_off + co) * a.Cin_total + a.ci_off + ci) * 27 + g * 3 + kx] = v16[c];
        }
      } else if (kx < 3 && ci < a.C) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int co = cob * 16 + c;
          if (co < a.Cout) pr[((int64_t)(a.co_off + co) * a.Cin_total + a.ci_off + ci) * 27 + g * 3 + kx] = 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// TMA-fed variant (W % 4 == 0, 16-byte aligned tensors): same GEMM, but the raw tiles arrive through rank-5 tensor maps
// (zero fill outside the volume and beyond the channel count, so the producers carry no bounds checks and no address
// arithmetic) and the tiles of a CTA walk along z, so each tile brings ONE new dY plane; the other two are still in
// the five-slot plane ring.  Warp 0 = MMA issue, warp 1 = TMA issue, warps 2..13 = B producers, 14..15 = A producers.
//   rawY slot: [16 co][6 rows y0-1..y0+4][16 x]   box {16, 6, 1, 16, 1}
//   rawX slot: [16 ci][4 rows y0..y0+3][24 x from x0-4]   box {24, 4, 1, 16, 1}   (inner coordinate stays 16-byte aligned)
constexpr int WV_YS = 5, WV_XS = 3;
constexpr int WV_RAW_BYTES = 16 * 6 * 16 * 4;  // 6144 for both kinds of slot
static_assert(16 * 4 * 24 * 4 == WV_RAW_BYTES, "raw slots share one size");
constexpr int WV_RAW_OFF = 2 * WU_STAGE_BYTES;  // 173056: multiple of 128
static_assert(WV_RAW_OFF % 128 == 0, "TMA destination alignment");
constexpr int WV_SMEM_BYTES = WV_RAW_OFF + (WV_YS + WV_XS) * WV_RAW_BYTES + 128;
constexpr int WV_NPROD = 448;

struct WgUmmaTmaArgs {
  float* partials;  // [region][P1+P2][H1+H2][27]
  int64_t region_stride;
  int H1, H2, P1, P2;   // channels of the halo-side tensors (x, or dy when transposed) and of the plain-side tensors
  int nH1, nP1, nPB;    // 16-channel blocks of the first halo tensor / first plain tensor, plain blocks in total
  int D;
  int tiles_x, tiles_y;
  int ncols, zlen, nunits;  // work unit = (column, z segment of zlen planes); unit u -> CTA u % gridDim.y
  float* bias_partials;     // nullable: [region][P1+P2] sums of the plain-side tensor (= bias gradient when that is dY)
  const float* amax_h;      // 3xFP16 kernel: device pointers to upper bounds of max|halo-side tensors| and
  const float* amax_p;      // max|plain-side tensors| (absmax_kernel), which fix the power-of-two scales
  int single_pass;          // 3xFP16 kernel: 1 = only the hi x hi MMA (reduced-precision mode)
  unsigned long long* dbg;  // optional cycle counters (DA_UMMA_DEBUG=1, 3xFP16 kernel): MMA warp waiting for operands / total,
                            // B producer warp waiting for raw tiles / for a free stage / total, TMA thread waiting, tiles
};

__global__ void __launch_bounds__(WU_THREADS, 1)
conv3d_wgrad_umma_tma_kernel(const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_h2,
                             const __grid_constant__ CUtensorMap map_p1, const __grid_constant__ CUtensorMap map_p2, WgUmmaTmaArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2], empty[2], rawfull[3], consumed[3], done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // channel blocks: halo side (rows of dW's inner index) and plain side, each possibly split over two tensors (the
  // concatenation is never materialised); a block never straddles the two
  int cob = blockIdx.x % a.nPB, cib = blockIdx.x / a.nPB;
  const bool h2 = cib >= a.nH1, p2 = cob >= a.nP1;
  if (h2) cib -= a.nH1;
  if (p2) cob -= a.nP1;
  const CUtensorMap* map_x = h2 ? &map_h2 : &map_h1;
  const CUtensorMap* map_dy = p2 ? &map_p2 : &map_p1;
  const int hC = h2 ? a.H2 : a.H1, pC = p2 ? a.P2 : a.P1;     // channels of the chosen tensors
  const int hg0 = (h2 ? a.H1 : 0), pg0 = (p2 ? a.P1 : 0);     // their offsets in the concatenated index
  const int region = blockIdx.y;
  // work units (column, z segment) are dealt round-robin: CTAs running together walk neighbouring columns in lockstep
  const int R = gridDim.y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], WV_NPROD); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&rawfull[i], 1); mbar_init(&consumed[i], WV_NPROD); }
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  float* rawY = reinterpret_cast<float*>(smem + WV_RAW_OFF);
  float* rawX = reinterpret_cast<float*>(smem + WV_RAW_OFF + WV_YS * WV_RAW_BYTES);

  if (warp == 0) {
    // =============================== MMA issue ===============================
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(WU_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc0 = umma_desc(0, 256, 128), bdesc0 = umma_desc(0, WU_BP, 128);
    const uint32_t base_s = smem_u32(smem);
    uint32_t first = 1;
    int ntl = 0;
    for (int u = region; u < a.nunits; u += R) { const int zb = (u / a.ncols) * a.zlen; ntl += min(a.D, zb + a.zlen) - zb; }
    for (int k = 0; k < ntl; ++k) {
      const int s = k & 1;
      mbar_wait(&full[s], (k >> 1) & 1);
      tc_fence_after();
      if (elected) {
        const uint32_t a_hi = base_s + (uint32_t)s * WU_STAGE_BYTES, a_lo = a_hi + WU_A_BYTES;
        const uint32_t b_hi = a_hi + 2 * WU_A_BYTES, b_lo = b_hi + WU_B_BYTES;
#pragma unroll
        for (int ks = 0; ks < WU_XT / 2; ++ks) {
          const uint32_t ao = (uint32_t)(2 * ks) * 256, bo = (uint32_t)(2 * ks) * WU_BP;
          umma_tf32(tmem, umma_desc_at(adesc0, a_hi + ao), umma_desc_at(bdesc0, b_hi + bo), idesc, first ? 0u : 1u);
          first = 0;
          umma_tf32(tmem, umma_desc_at(adesc0, a_lo + ao), umma_desc_at(bdesc0, b_hi + bo), idesc, 1u);
          umma_tf32(tmem, umma_desc_at(adesc0, a_hi + ao), umma_desc_at(bdesc0, b_lo + bo), idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elected) umma_commit(&done);
    __syncwarp();
  } else if (warp == 1) {
    // =============================== TMA issue ===============================
    if (lane == 0) {
      tma_prefetch_desc(map_x);
      tma_prefetch_desc(map_dy);
      int q = -1;  // sequence number of the newest dY plane in the ring
      int k = 0;
      for (int u = region; u < a.nunits; u += R) {
        int col = u % a.ncols;
        const int zb = (u / a.ncols) * a.zlen, ze = min(a.D, zb + a.zlen);
        const int bx = col % a.tiles_x; col /= a.tiles_x;
        const int by = col % a.tiles_y;
        const int n = col / a.tiles_y;
        const int x0 = bx * WU_XT, y0 = by * WU_YT;
        for (int z = zb; z < ze; ++z, ++k) {
        const bool fresh = z == zb;  // new unit: all three planes; otherwise only plane z+1
        // slots about to be overwritten were last read by tile k-1 (fresh) or k-3 (steady state)
        const int dep = fresh ? k - 1 : k - 3;
        if (dep >= 0) mbar_wait(&consumed[dep % 3], (dep / 3) & 1);
        uint64_t* bar = &rawfull[k % 3];
        mbar_expect_tx(bar, (uint32_t)WV_RAW_BYTES * (fresh ? 4u : 2u));
        tma_load_5d(rawX + (k % WV_XS) * (WV_RAW_BYTES / 4), map_x, bar, x0 - 4, y0, z, cib * 16, n);
        if (fresh) {
          tma_load_5d(rawY + ((q + 1) % WV_YS) * (WV_RAW_BYTES / 4), map_dy, bar, x0, y0 - 1, z - 1, cob * 16, n);
          tma_load_5d(rawY + ((q + 2) % WV_YS) * (WV_RAW_BYTES / 4), map_dy, bar, x0, y0 - 1, z, cob * 16, n);
          q += 3;
        } else {
          q += 1;
        }
        tma_load_5d(rawY + (q % WV_YS) * (WV_RAW_BYTES / 4), map_dy, bar, x0, y0 - 1, z + 1, cob * 16, n);
        }
      }
    }
    __syncwarp();
  } else if (warp < 14) {
    // =============================== B producers: (x, co, kz) fixed per (thread, j) ===============================
    const int tp = threadIdx.x - 64;
    const int tx = tp & 15;
    int tkz[2], tco[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tb = tp + 384 * j;
      tco[j] = (tb >> 4) & 15; tkz[j] = tb >> 8;
    }
    int q = -1, k = 0;
    double bsum = 0.0;  // bias gradient: the kz = 1 task holds every dY value of the tile exactly once (rows 1..4); fp64 across tiles
    for (int u = region; u < a.nunits; u += R) {
      const int zb = (u / a.ncols) * a.zlen, ze = min(a.D, zb + a.zlen);
      for (int z = zb; z < ze; ++z, ++k) {
      const int s = k & 1, use = k >> 1;
      q += (z == zb) ? 3 : 1;
      mbar_wait(&rawfull[k % 3], (k / 3) & 1);
      float v[2][6];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float* p = rawY + ((q - tkz[j]) % WV_YS) * (WV_RAW_BYTES / 4) + tco[j] * 96 + tx;  // plane z - kz + 1
#pragma unroll
        for (int r = 0; r < 6; ++r) v[j][r] = p[r * 16];
        if (tkz[j] == 1) bsum += (double)((v[j][1] + v[j][2]) + (v[j][3] + v[j][4]));
      }
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
      uint8_t* st = smem + s * WU_STAGE_BYTES + 2 * WU_A_BYTES;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {  // dY row y0 + e - ky + 1  =  v[e - ky + 2]
            const float val = v[j][e - ky + 2];
            h[e] = __uint_as_float(__float_as_uint(val) & 0xffffe000u);
            l[e] = val - h[e];
          }
          const int off = tx * WU_BP + ((tkz[j] * 3 + ky) * 16 + tco[j]) * 16;
          *reinterpret_cast<float4*>(st + off) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(st + WU_B_BYTES + off) = make_float4(l[0], l[1], l[2], l[3]);
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
      mbar_arrive(&consumed[k % 3]);
      }
    }
    if (a.bias_partials && !h2 && cib == 0) {
      // the 16 x-lanes of one channel sit in one half warp; threads 128..255 of the producer set hold no kz = 1 task
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
      const int co = (tp < 128) ? tco[1] : tco[0];
      if (tx == 0 && (tp < 128 || tp >= 256) && cob * 16 + co < pC)
        a.bias_partials[(int64_t)region * (a.P1 + a.P2) + pg0 + cob * 16 + co] = (float)bsum;
    }
  } else {
    // =============================== A producers: (xc, ci) fixed per (thread, j) ===============================
    // lanes 0..7 of every quarter warp hold eight different ci (16 B apart in the tile): conflict-free 16-byte stores
    const int w = warp - 14;
    int txc[6], tci[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int g = w + 2 * j;  // 12 groups of 32 tasks: (ci half, four chunks)
      tci[j] = (lane & 7) + 8 * (g & 1);
      txc[j] = (lane >> 3) + 4 * (g >> 1);
    }
    int ntl = 0;
    for (int u = region; u < a.nunits; u += R) { const int zb = (u / a.ncols) * a.zlen; ntl += min(a.D, zb + a.zlen) - zb; }
    for (int k = 0; k < ntl; ++k) {
      const int s = k & 1, use = k >> 1;
      mbar_wait(&rawfull[k % 3], (k / 3) & 1);
      const float* xs = rawX + (k % WV_XS) * (WV_RAW_BYTES / 4);
      float v[6][4];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        // chunk xc holds input x = x0 - 1 + xc = raw column xc + 3; chunks past 20 feed only the unused rows t >= 3
        const bool ok = txc[j] + 3 < 24;
        const float* p = xs + tci[j] * 96 + txc[j] + 3;
#pragma unroll
        for (int r = 0; r < 4; ++r) v[j][r] = ok ? p[r * 24] : 0.f;
      }
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
      uint8_t* st = smem + s * WU_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h[e] = __uint_as_float(__float_as_uint(v[j][e]) & 0xffffe000u);
          l[e] = v[j][e] - h[e];
        }
        const int off = (txc[j] * 16 + tci[j]) * 16;
        *reinterpret_cast<float4*>(st + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(st + WU_A_BYTES + off) = make_float4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
      mbar_arrive(&consumed[k % 3]);
    }
  }

  // ---- final read of the accumulator: rows 16*t + ci (t = kx < 3), columns (kz*3+ky)*16 + co ----
  if (warp == 4 || warp == 5) {
    mbar_wait(&done, 0);
    tc_fence_after();
    const int r = (warp - 4) * 32 + lane;
    const int kx = r >> 4, ci = cib * 16 + (r & 15);
    float* pr = a.partials + (int64_t)region * a.region_stride;
#pragma unroll 1
    for (int g = 0; g < 9; ++g) {
      float v16[16];
      tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 16), v16);
      tmem_ld_wait();
      if (kx < 3 && ci < hC) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int co = cob * 16 + c;
          if (co < pC) pr[((int64_t)(pg0 + co) * (a.H1 + a.H2) + hg0 + ci) * 27 + g * 3 + kx] = v16[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// 3xFP16 variant of the TMA-fed kernel (the default): kind::f16 MMAs on fp16 hi/lo pairs of the operands scaled per
// tensor by a power of two (see conv3d_umma.inc.cuh, mode 1) contract K = 16 positions per instruction at the cycle cost
// of a K = 8 tf32 one (tools/probes/umma16_probe.cu), and the halo-side block may be 32 channels wide, so 96 of the 128
// MMA rows are useful instead of 48.
//   D[(t, ci)][(kz, ky, co)] += sum_k A[(t, ci)][k] * B[(kz, ky, co)][k]     k = (pair p, xe, ye): positions x0+2p+xe, y0+ye
// * A 16-byte K chunk holds 8 halves = one x-PAIR x 4 y-rows.  The kx shift is again free, but now through two
//   interleaved images of X: block b = [CIB ci][(xe, ye)] holds the positions x0-1+b+xe, i.e. even blocks are the
//   pairs of the kx = 0 / 2 alignment and odd blocks those of kx = 1.  MMA row CIB*t + ci of the K chunk of pair p
//   then simply reads block 2p + t (LBO = two blocks = one pair, SBO = 128 B; rows t >= 3 compute garbage).
// * B as before: each dY pair is replicated into the nine (kz, ky) row groups, pitch per pair 144*16 + 16 B.
// * 12 MMAs per 16 x 4 tile (four K steps x hi*hi, lo*hi, hi*lo), two operand stages, raw tiles by TMA with the
//   channel dimension in the MIDDLE of the box ([row][channel][x]): the producers' lanes run along channels (A, rows of
//   28 floats) or x-pairs (B, rows of 16 floats) without bank conflicts.
// Warp roles: 0 = MMA issue, 1 = TMA issue, 2..13 = B producers, 14.. = A producers (2 warps at CIB 16, 4 at CIB 32).
constexpr int WB_YS = 7, WB_XS = 4, WB_NR = 4;   // raw dY plane ring, raw X tile ring, tiles in flight (barrier ring)
constexpr int WB_XBOX = 28;                     // raw X row: x0-4 .. x0+23 (112 B; 28-float channel pitch = 8 distinct banks)
constexpr int WB_RAWY_BYTES = 6 * 16 * 16 * 4;  // [6 rows][16 co][16 x]
template <int CIB>
struct WbCfg {
  static constexpr int NAW = CIB / 8;                          // A producer warps
  static constexpr int NPROD = 384 + 32 * NAW;                 // producer threads
  static constexpr int THREADS = 64 + NPROD;
  static constexpr int NBLK = CIB == 32 ? 18 : 22;             // A blocks an MMA may touch: 2*7 + (128/CIB - 1) + 1
  static constexpr int BLK_BYTES = CIB * 16;
  static constexpr int A_BYTES = NBLK * BLK_BYTES;             // per hi / lo half
  static constexpr int B_BYTES = 8 * WU_BP;                    // 8 pairs
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int RAWX_BYTES = 4 * CIB * WB_XBOX * 4;     // [4 rows][CIB ci][28 x]
  static constexpr int RAWY_OFF = 2 * STAGE_BYTES;
  static constexpr int RAWX_OFF = RAWY_OFF + WB_YS * WB_RAWY_BYTES;
  static constexpr int SMEM_BYTES = RAWX_OFF + WB_XS * RAWX_BYTES + 128;
  static_assert(STAGE_BYTES % 128 == 0 && RAWX_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(SMEM_BYTES <= 232448 - 1024, "shared memory");
};

template <int CIB, bool DBG>
__global__ void __launch_bounds__(WbCfg<CIB>::THREADS, 1)
conv3d_wgrad_umma16_kernel(const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_h2,
                           const __grid_constant__ CUtensorMap map_p1, const __grid_constant__ CUtensorMap map_p2, WgUmmaTmaArgs a) {
  using Cfg = WbCfg<CIB>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2], empty[2], rawfull[WB_NR], consumed[WB_NR], done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // channel blocks: halo side in blocks of CIB (a.nH1 = blocks of the first halo tensor), plain side in blocks of 16
  int cob = blockIdx.x % a.nPB, cib = blockIdx.x / a.nPB;
  const bool h2 = cib >= a.nH1, p2 = cob >= a.nP1;
  if (h2) cib -= a.nH1;
  if (p2) cob -= a.nP1;
  const CUtensorMap* map_x = h2 ? &map_h2 : &map_h1;
  const CUtensorMap* map_dy = p2 ? &map_p2 : &map_p1;
  const int hC = h2 ? a.H2 : a.H1, pC = p2 ? a.P2 : a.P1;
  const int hg0 = (h2 ? a.H1 : 0), pg0 = (p2 ? a.P1 : 0);
  const int region = blockIdx.y;
  const int R = gridDim.y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], Cfg::NPROD); mbar_init(&empty[i], 1); }
    for (int i = 0; i < WB_NR; ++i) { mbar_init(&rawfull[i], 1); mbar_init(&consumed[i], Cfg::NPROD); }
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  float* rawY = reinterpret_cast<float*>(smem + Cfg::RAWY_OFF);
  float* rawX = reinterpret_cast<float*>(smem + Cfg::RAWX_OFF);
  int ntl = 0;   // tiles this CTA walks
  for (int u = region; u < a.nunits; u += R) { const int zb = (u / a.ncols) * a.zlen; ntl += min(a.D, zb + a.zlen) - zb; }

  if (warp == 0) {
    // =============================== MMA issue ===============================
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const uint32_t idesc = umma_idesc<true>(128, WU_N);
    const uint64_t adesc0 = umma_desc(0, 2 * Cfg::BLK_BYTES, 128), bdesc0 = umma_desc(0, WU_BP, 128);
    const uint32_t base_s = smem_u32(smem);
    uint32_t first = 1;
    long long d_wait = 0;
    const long long d_begin = DBG ? clock64() : 0;
    for (int k = 0; k < ntl; ++k) {
      const int s = k & 1;
      const long long d0 = DBG ? clock64() : 0;
      mbar_wait(&full[s], (k >> 1) & 1);
      if (DBG) d_wait += clock64() - d0;
      tc_fence_after();
      if (elected) {
        const uint32_t a_hi = base_s + (uint32_t)s * Cfg::STAGE_BYTES, a_lo = a_hi + Cfg::A_BYTES;
        const uint32_t b_hi = a_hi + 2 * Cfg::A_BYTES, b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // K step = pairs 2q, 2q+1 = A blocks 4q.. (+2 per pair), B pairs 2q, 2q+1
          const uint32_t ao = (uint32_t)(4 * q) * Cfg::BLK_BYTES, bo = (uint32_t)(2 * q) * WU_BP;
          umma_f16(tmem, umma_desc_at(adesc0, a_hi + ao), umma_desc_at(bdesc0, b_hi + bo), idesc, first ? 0u : 1u);
          first = 0;
          if (!a.single_pass) {
            umma_f16(tmem, umma_desc_at(adesc0, a_lo + ao), umma_desc_at(bdesc0, b_hi + bo), idesc, 1u);
            umma_f16(tmem, umma_desc_at(adesc0, a_hi + ao), umma_desc_at(bdesc0, b_lo + bo), idesc, 1u);
          }
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elected) umma_commit(&done);
    __syncwarp();
    if (DBG && a.dbg && lane == 0) {
      atomicAdd(a.dbg + 0, (unsigned long long)d_wait); atomicAdd(a.dbg + 1, (unsigned long long)(clock64() - d_begin));
      atomicAdd(a.dbg + 6, (unsigned long long)ntl); atomicAdd(a.dbg + 7, 1ull);
    }
  } else if (warp == 1) {
    // =============================== TMA issue ===============================
    if (lane == 0) {
      tma_prefetch_desc(map_x);
      tma_prefetch_desc(map_dy);
      int q = -1;  // sequence number of the newest dY plane in the ring
      int k = 0;
      long long d_wait = 0;
      for (int u = region; u < a.nunits; u += R) {
        int col = u % a.ncols;
        const int zb = (u / a.ncols) * a.zlen, ze = min(a.D, zb + a.zlen);
        const int bx = col % a.tiles_x; col /= a.tiles_x;
        const int by = col % a.tiles_y;
        const int n = col / a.tiles_y;
        const int x0 = bx * WU_XT, y0 = by * WU_YT;
        for (int z = zb; z < ze; ++z, ++k) {
          const bool fresh = z == zb;  // new unit: all three planes; otherwise only plane z+1
          // ring slots about to be overwritten: the X slot and barrier of tile k-4, dY planes last read by tile k-5
          // (steady state) or by tile k-2 at the latest (new unit: three planes, sequence numbers jump by three)
          const int dep = fresh ? k - 2 : k - WB_NR;
          const long long d0 = DBG ? clock64() : 0;
          if (dep >= 0) mbar_wait(&consumed[dep % WB_NR], (dep / WB_NR) & 1);
          if (DBG) d_wait += clock64() - d0;
          uint64_t* bar = &rawfull[k % WB_NR];
          mbar_expect_tx(bar, (uint32_t)Cfg::RAWX_BYTES + (uint32_t)WB_RAWY_BYTES * (fresh ? 3u : 1u));
          tma_load_5d(rawX + (k % WB_XS) * (Cfg::RAWX_BYTES / 4), map_x, bar, x0 - 4, cib * CIB, y0, z, n);
          if (fresh) {
            tma_load_5d(rawY + ((q + 1) % WB_YS) * (WB_RAWY_BYTES / 4), map_dy, bar, x0, cob * 16, y0 - 1, z - 1, n);
            tma_load_5d(rawY + ((q + 2) % WB_YS) * (WB_RAWY_BYTES / 4), map_dy, bar, x0, cob * 16, y0 - 1, z, n);
            q += 3;
          } else {
            q += 1;
          }
          tma_load_5d(rawY + (q % WB_YS) * (WB_RAWY_BYTES / 4), map_dy, bar, x0, cob * 16, y0 - 1, z + 1, n);
        }
      }
      if (DBG && a.dbg) atomicAdd(a.dbg + 5, (unsigned long long)d_wait);
    }
    __syncwarp();
  } else if (warp < 14) {
    // =============================== B producers: task = (pair p, co, kz) ===============================
    const int tb = threadIdx.x - 64;
    const int p = tb & 7, co = (tb >> 3) & 15, kz = tb >> 7;
    int q = -1, k = 0;
    // bias gradient: the kz = 1 tasks see every dY value of the tile exactly once (rows 1..4).  The gradient of a
    // translation-invariant loss sums to ~0 over the volume (flow.bias): fp64 across the thread's tiles keeps the
    // cancellation from amplifying fp32 round-off (one DADD per tile)
    double bsum = 0.0;
    const float sc = pow2f(scale_exp_from_amax(__ldg(a.amax_p)));
    long long d_raw = 0, d_empty = 0;
    const long long d_begin = DBG ? clock64() : 0;
    for (int u = region; u < a.nunits; u += R) {
      const int zb = (u / a.ncols) * a.zlen, ze = min(a.D, zb + a.zlen);
      for (int z = zb; z < ze; ++z, ++k) {
        const int s = k & 1, use = k >> 1;
        q += (z == zb) ? 3 : 1;
        const long long d0 = DBG ? clock64() : 0;
        mbar_wait(&rawfull[k % WB_NR], (k / WB_NR) & 1);
        if (DBG) d_raw += clock64() - d0;
        const float* src = rawY + ((q - kz) % WB_YS) * (WB_RAWY_BYTES / 4) + co * 16 + 2 * p;  // plane z - kz + 1
        float2 v[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) v[r] = *reinterpret_cast<const float2*>(src + r * 256);
        if (kz == 1) bsum += (double)(((v[1].x + v[1].y) + (v[2].x + v[2].y)) + ((v[3].x + v[3].y) + (v[4].x + v[4].y)));
#pragma unroll
        for (int r = 0; r < 6; ++r) { v[r].x *= sc; v[r].y *= sc; }
        const long long d1 = DBG ? clock64() : 0;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        if (DBG) d_empty += clock64() - d1;
        uint8_t* st = smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + p * WU_BP + (kz * 3 * 16 + co) * 16;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {   // K element (xe, ye) = dY row y0 + ye - ky + 1 = v[ye - ky + 2]
          uint4 h, l;
          split_f16x2(v[2 - ky].x, v[3 - ky].x, h.x, l.x);
          split_f16x2(v[4 - ky].x, v[5 - ky].x, h.y, l.y);
          split_f16x2(v[2 - ky].y, v[3 - ky].y, h.z, l.z);
          split_f16x2(v[4 - ky].y, v[5 - ky].y, h.w, l.w);
          *reinterpret_cast<uint4*>(st + ky * 256) = h;
          *reinterpret_cast<uint4*>(st + Cfg::B_BYTES + ky * 256) = l;
        }
        fence_proxy_async();
        mbar_arrive(&full[s]);
        mbar_arrive(&consumed[k % WB_NR]);
      }
    }
    if (DBG && a.dbg && tb == 0) {
      atomicAdd(a.dbg + 2, (unsigned long long)d_raw); atomicAdd(a.dbg + 3, (unsigned long long)d_empty);
      atomicAdd(a.dbg + 4, (unsigned long long)(clock64() - d_begin));
    }
    if (a.bias_partials && !h2 && cib == 0 && kz == 1) {
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 4);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
      if (p == 0 && cob * 16 + co < pC) a.bias_partials[(int64_t)region * (a.P1 + a.P2) + pg0 + cob * 16 + co] = (float)bsum;
    }
  } else {
    // =============================== A producers: task = (block b, ci) ===============================
    // lanes run along ci: conflict-free 16-byte stores, and the 28-float channel pitch of the raw tile spreads the
    // 4-byte reads over 8 banks
    constexpr int NT = 17 * CIB;                       // blocks 0..16 carry data (rows t <= 2 of pairs 0..7)
    constexpr int TPT = (NT + 32 * Cfg::NAW - 1) / (32 * Cfg::NAW);
    const int ta0 = (warp - 14) * 32 + lane;
    const float sc = pow2f(scale_exp_from_amax(__ldg(a.amax_h)));
    for (int k = 0; k < ntl; ++k) {
      const int s = k & 1, use = k >> 1;
      mbar_wait(&rawfull[k % WB_NR], (k / WB_NR) & 1);
      const float* xs = rawX + (k % WB_XS) * (Cfg::RAWX_BYTES / 4);
      float v[TPT][8];
#pragma unroll
      for (int j = 0; j < TPT; ++j) {
        const int ta = ta0 + 32 * Cfg::NAW * j;
        const int ci = ta % CIB, b = ta / CIB;
        // block b, element (xe, ye) = position x0 - 1 + b + xe, row y0 + ye = raw column 3 + b + xe
        const float* src = xs + ci * WB_XBOX + 3 + b;
#pragma unroll
        for (int ye = 0; ye < 4; ++ye) {
          v[j][ye] = ta < NT ? src[ye * (CIB * WB_XBOX)] : 0.f;
          v[j][4 + ye] = ta < NT ? src[ye * (CIB * WB_XBOX) + 1] : 0.f;
        }
      }
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
      uint8_t* st = smem + s * Cfg::STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < TPT; ++j) {
        const int ta = ta0 + 32 * Cfg::NAW * j;
        if (ta < NT) {
          uint4 h, l;
          split_f16x2(v[j][0] * sc, v[j][1] * sc, h.x, l.x);
          split_f16x2(v[j][2] * sc, v[j][3] * sc, h.y, l.y);
          split_f16x2(v[j][4] * sc, v[j][5] * sc, h.z, l.z);
          split_f16x2(v[j][6] * sc, v[j][7] * sc, h.w, l.w);
          *reinterpret_cast<uint4*>(st + ta * 16) = h;               // (b * CIB + ci) * 16
          *reinterpret_cast<uint4*>(st + Cfg::A_BYTES + ta * 16) = l;
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
      mbar_arrive(&consumed[k % WB_NR]);
    }
  }

  // ---- final read of the accumulator: rows CIB*t + ci (t = kx < 3), columns (kz*3+ky)*16 + co ----
  if (warp >= 4 && warp < 4 + (3 * CIB + 31) / 32) {
    mbar_wait(&done, 0);
    tc_fence_after();
    const int r = (warp - 4) * 32 + lane;   // TMEM lane = accumulator row; warp 4 + i reads lane quadrant i
    const int kx = r / CIB, ci = cib * CIB + r % CIB;
    float* pr = a.partials + (int64_t)region * a.region_stride;
    const float us = pow2f(max(-126, -(scale_exp_from_amax(__ldg(a.amax_h)) + scale_exp_from_amax(__ldg(a.amax_p)))));
#pragma unroll 1
    for (int g = 0; g < 9; ++g) {
      float v16[16];
      tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 16), v16);
      tmem_ld_wait();
      if (kx < 3 && ci < hC) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int co = cob * 16 + c;
          if (co < pC) pr[((int64_t)(pg0 + co) * (a.H1 + a.H2) + hg0 + ci) * 27 + g * 3 + kx] = ntl > 0 ? v16[c] * us : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}
