// Tensor-core (tcgen05) variant of the k3 s1 p1 convolution forward / data gradient.  Included by conv3d.cu inside its
// anonymous namespace.  Three operand formats (template MODE), all with fp32 accumulation in TMEM:
//   0  3xTF32: kind::tf32, 16 input channels per launch (four 16-byte K chunks of 4 floats), products exact to 2^-21
//   1  3xFP16: kind::f16 on fp16 hi/lo pairs of the operands scaled by a power of two (per tensor, from its max-abs:
//      max|x| * 2^k in [2^13, 2^14), so hi + lo carries 22 significant bits for everything within 2^17 of the maximum
//      and an absolute error below 2^-39 max|x| beneath that -- the accuracy class of 3xTF32, measured against fp64
//      in the parity tests), 32 input channels per launch (four 16-byte K chunks of 8 halves): same shared-memory
//      bytes, same MMA count and the same ~104 cycles per MMA as mode 0 (tools/probes/umma16_probe.cu), i.e. twice
//      the channels per MMA.  (bf16 pairs need no scaling but carry 16 bits: weight gradients of the first layers
//      came out 1e-2 off the fp64 oracle, against 1e-4 for 3xTF32 -- measured, rejected.)
//   2  3xFP16 with 16 input channels per launch (two K chunks): narrow layers and the remainder of Cin % 32
//   3, 4  modes 1, 2 with the input planes staged by TMA: the register-staged producers of modes 0-2 hold 56 loads per
//      thread in flight = 21 KB per SM, and one DRAM round trip per batch made them -- not the tensor pipe -- the
//      limit of a plane step (4.1k cycles per 16 channels, measured with the cycle counters).  Here a plane is cut
//      into 16-channel units; raw fp32 units ([16 ch][8 rows][48 x] boxes, zero-filled outside the volume) arrive
//      through a ring of 3-4 TMA slots (49-74 KB in flight, no registers), the three producer warps only convert
//      shared memory -> scaled fp16 hi/lo -> shared memory, and the MMAs of a plane run unit by unit (the second unit
//      accumulates into the same TMEM blocks).  Needs W % 4 == 0, 16-byte aligned tensors and a concatenation boundary
//      on a multiple of 8 channels; otherwise modes 1, 2 serve.
//
// Implicit GEMM, output-stationary in TMEM:
//   D[f][(kx,co)] = sum_{kz,ky,ci} X[ci][plane zo+kz-1][f + (ky-1)*PX] * W[co][ci][kz][ky][kx]      f = in-plane position
//   Y[co][q]      = D[q-1][(0,co)] + D[q][(1,co)] + D[q+1][(2,co)]                                 (kx fold in the epilogue)
// * M = 128 positions per MMA: consecutive positions of a (TY+2) x PX plane tile flattened row-major, so a (ky) shift
//   is just a start-address offset of +-PX rows in the A descriptor (un-swizzled K-major canonical layout
//   [ci/4][position][4 floats]; measured to work, tools/probes/umma_probe.cu) and kz selects one of three ring slots.
// * N = 144 = 3 planes x 3 kx x 16 co.  One tcgen05.mma costs >= 97 cycles regardless of N <= 128 (shared-memory A-read
//   floor, same probe), so as many taps as possible are folded into N: kx (resolved in the epilogue) and kz.  The kz fold
//   makes the accumulator three rotating 48-column blocks: input plane p adds its kz = 0/1/2 contribution to the blocks
//   of output planes p+1, p, p-1 with ONE instruction, because the weight rows are stored in kz order 2,1,0,2,1 and the
//   B descriptor starts at the rotation the step needs.  After step p the block of output plane p-1 is complete: the
//   epilogue reads it, zeroes it (tcgen05.st) and hands it back.  36 MMAs per plane instead of 108.
// * fp32 accuracy from three TF32 MMAs per K step: A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with hi = x & 0xffffe000 (the unit
//   truncates), lo = x - hi, both computed by the producer warps while they restage the planar input.
// * z streaming: a CTA owns a TY x TX column and walks ZG output planes; each input plane is staged once (3-slot ring,
//   only the current plane is read by a step, so the producers run two planes ahead).
// Warp roles: 0-7 epilogue (M tile w/4, TMEM lane quadrant w%4), 8 MMA issue + TMEM allocation, 9-16 producers.
// Channel blocking: 16 output x 16 input channels per launch; further input-channel chunks accumulate into the output
// in global memory (bias / activation applied by the last chunk), further output blocks are separate launches.

constexpr int UM_TX = 40, UM_TY = 6, UM_PX = UM_TX + 2;  // (a 128-byte aligned pitch of 48 was measured: no gain)
constexpr int UM_PLANE = (UM_TY + 2) * UM_PX;  // 336 positions staged per plane
constexpr int UM_PFA = 344;                    // allocated positions per channel chunk (>= 2*128 + 2*PX)
constexpr int UM_MT = 2;                       // M tiles per plane
constexpr int UM_MSTEP = 126;                  // the two M tiles overlap by two rows: each folds kx on its own (rows 1..126)
constexpr int UM_NPROD = 96;                   // producer threads (3 warps): 20 warps in all = 5 per SM sub-partition, which leaves 96
                                               // registers per thread (6 per sub-partition cap them at 80 and the epilogue spills)
constexpr int UM_NEPI = 512;                   // epilogue threads: warp w handles TMEM lane quadrant w%4, M tile (w/4)%2, output-channel half w/8
constexpr int UM_THREADS = UM_NEPI + 32 + UM_NPROD;  // warps 0-15 epilogue, 16 MMA issue, 17-19 producers (640 threads)
constexpr int UM_MMA_WARP = UM_NEPI / 32;
constexpr int UM_CB = 16;                      // output channels per launch
constexpr int UM_NB = 3 * UM_CB;               // columns of one accumulator block: (kx, co)
constexpr int UM_N = 3 * UM_NB;                // MMA N: three blocks = three output planes in flight
constexpr int UM_WROWS = 5 * UM_NB;            // weight rows: kz blocks in the order 2,1,0,2,1 (any cyclic rotation is contiguous)
static_assert(UM_MT * 128 + 2 * UM_PX <= UM_PFA, "shifted A rows must stay inside the slot");
static_assert(UM_MT * UM_MSTEP >= UM_TY * UM_PX, "M tiles must cover the output rows");

constexpr int UM_RAWX = 48;                    // TMA modes: raw row = x0-4 .. x0+43 (the inner box coordinate must be 16-byte aligned)
template <int MODE>
struct UmmaCfg {
  static constexpr bool BF = MODE != 0;
  static constexpr bool TMAIN = MODE >= 3;
  static constexpr int EPC = BF ? 8 : 4;                             // input channels per 16-byte K chunk
  static constexpr int NCH = (MODE == 2 || MODE == 4) ? 2 : 4;       // K chunks per launch
  static constexpr int KC = EPC * NCH;                               // input channels per launch: 16, 32, 16, 32, 16
  static constexpr int SCH = TMAIN ? 2 : NCH;                        // K chunks per ring slot (TMA modes: one 16-channel unit)
  static constexpr int NSUB = NCH / SCH;                             // units per plane
#ifndef DA_M3_RING
#define DA_M3_RING 2
#define DA_M3_RAW 3
#endif
  static constexpr int NRING = (MODE == 3) ? DA_M3_RING : 3;
  static constexpr int SLOT_BYTES = 2 * SCH * UM_PFA * 16;           // hi + lo
  static constexpr int RING_BYTES = NRING * SLOT_BYTES;
  static constexpr int NRAW = TMAIN ? (MODE == 3 ? DA_M3_RAW : 4) : 0;       // raw fp32 units in flight
  static constexpr int RAW_SLOT_BYTES = 16 * (UM_TY + 2) * UM_RAWX * 4;   // [16 ch][8 rows][48 x]
  static constexpr int RAW_BYTES = NRAW * RAW_SLOT_BYTES;
  static constexpr int W_BYTES = 2 * 3 * NCH * UM_WROWS * 16;        // [hi|lo][ky][ci/EPC][row][16 bytes]
  static constexpr int SMEM_BYTES = RAW_BYTES + RING_BYTES + W_BYTES + 128;   // raw slots first (TMA destinations: 128-byte aligned)
  static constexpr int TMEM_COLS = 512;
  static_assert(UM_MT * UM_N <= 512, "accumulators exceed TMEM");
  static_assert(RAW_SLOT_BYTES % 128 == 0 && SLOT_BYTES % 128 == 0, "alignment");
  static_assert(SMEM_BYTES <= 232448 - 4096, "shared memory (static arrays take ~2.5 KB)");
};
constexpr int UMMA_IMG_STRIDE_BYTES = UmmaCfg<0>::W_BYTES;  // 92160: every weight image of a layer sits at this pitch

struct UmmaArgs {
  const float* x1; const float* x2; int C1, C2;
  const float* wimg;          // prepared weight image of (co block 0, this channel chunk); block ib is img_stride floats further
                              // (fp32 hi/lo in mode 0, scaled fp16 hi/lo otherwise)
  const float* amax_x;        // modes 1-4: device pointers to an upper bound of max|input| and to max|weight| (absmax_kernel),
  const float* amax_w;        // which fix the power-of-two scales of the two operands
  int single_pass;            // 1: only the hi x hi MMA (half-precision operands, fp32 accumulate: the reduced-precision mode)
  int64_t img_stride; int nco; // output-channel blocks of the layer = blockIdx.z % nco
  const float* bias; float* out;
  int N, D, H, W, Cout;
  int c0;                     // first input channel of this chunk
  int accumulate, last;       // add to the existing output; apply bias + activation
  int act; float slope;
  int tiles_x, tiles_y, zg;   // z planes per CTA
  int flags;                  // debug (DA_UMMA_FLAGS), unused at present
  unsigned long long* dbg;    // optional cycle counters of the MMA warp (DA_UMMA_DEBUG=1): acc wait, plane wait, issue, total, steps
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// descriptor with the constant fields (LBO, SBO, version) pre-assembled: only the 14-bit start address changes per MMA
__device__ __forceinline__ uint64_t umma_desc_at(uint64_t base_no_addr, uint32_t saddr) {
  return base_no_addr | (uint64_t)((saddr >> 4) & 0x3fff);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 covers fp16 and bf16 operands (the instruction descriptor says which); K = 16 per instruction
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool BF>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (BF) umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}
// instruction descriptor: fp32 accumulate, A and B both K-major, format 2 = tf32 / 0 = fp16, N at bit 17, M at bit 24
template <bool BF>
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
  const uint32_t fmt = BF ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Power-of-two scale 2^k that puts a tensor's max-abs into [2^13, 2^14) (fp16 overflows at 2^16), and its inverse.
// amax = 0 or denormal gives the clamp 2^100 (the products are zero anyway).
__device__ __forceinline__ int scale_exp_from_amax(float amax) {
  const int e = (int)((__float_as_uint(amax) >> 23) & 0xffu) - 127;   // floor(log2(amax)) of a normal number
  return max(-100, min(100, 13 - e));
}
__device__ __forceinline__ float pow2f(int k) { return __uint_as_float((uint32_t)(k + 127) << 23); }
// x (already scaled) = hi + lo with both halves fp16, round to nearest: |x - hi - lo| <= 2^-24 |x| while lo is a normal
// number, 2^-25 absolute below.  Packs two values per register, the first in the low half (= the lower K index).
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const __half2 l = __floats2half2_rn(a - __low2float(h), b - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// max|x| over up to four segments into out[slot] (slots hold non-negative floats: their bit patterns order like
// unsigned integers, so atomicMax works on them).  out must be zeroed before the launch.
struct AbsmaxArgs { const float* p[4]; int64_t n[4]; int slot[4]; };
__global__ void __launch_bounds__(256) absmax_kernel(AbsmaxArgs a, float* __restrict__ out) {
  __shared__ float red[8];
  for (int s = 0; s < 4; ++s) {
    if (!a.p[s] || a.n[s] <= 0) continue;   // uniform over the grid
    const float* p = a.p[s];
    const int64_t n = a.n[s];
    float m = 0.f;
    const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x, nth = (int64_t)gridDim.x * 256;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const int64_t n4 = n >> 2;
      for (int64_t i = tid; i < n4; i += nth) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
      for (int64_t i = (n4 << 2) + tid; i < n; i += nth) m = fmaxf(m, fabsf(__ldg(p + i)));
    } else {
      for (int64_t i = tid; i < n; i += nth) m = fmaxf(m, fabsf(__ldg(p + i)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float r = red[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) r = fmaxf(r, red[w]);
      atomicMax(reinterpret_cast<unsigned int*>(out) + a.slot[s], __float_as_uint(r));
    }
  }
}
// shared-space accesses with 32-bit addresses (generic pointers cost a register pair per dynamic offset)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

template <bool DBG, bool ACC, int MODE>   // cycle counters (DA_UMMA_DEBUG=1); a.accumulate (further input-channel chunks); operand format
__global__ void __launch_bounds__(UM_THREADS, 1)
conv3d_umma_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2, UmmaArgs a) {
  using Cfg = UmmaCfg<MODE>;
  constexpr int N = UM_N, NCH = Cfg::NCH, NB = UM_NB;
  constexpr bool BF = Cfg::BF;
  extern __shared__ uint8_t smem_raw[];
  // plane_full / plane_empty: ring slots (one plane, or one 16-channel unit of a plane in the TMA modes)
  __shared__ __align__(8) uint64_t plane_full[3], plane_empty[3], acc_full[UM_MT], acc_empty[UM_MT], raw_full[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ float edge_s[2][8][2][16];  // warp-edge rows of the kx fold (static: keeps LDS/STS addressing)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* rawbuf = smem;
  uint8_t* ring = smem + Cfg::RAW_BYTES;
  uint8_t* sw = ring + Cfg::RING_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tb = blockIdx.x;
  const int bx = tb % a.tiles_x, by = tb / a.tiles_x;
  const int X0 = bx * UM_TX, Y0 = by * UM_TY;
  const int z0 = blockIdx.y * a.zg;
  const int zcount = min(a.zg, a.D - z0);
  const int nsteps = zcount + 2;  // input planes z0-1 .. z0+zcount
  const int n = blockIdx.z / a.nco, ib = blockIdx.z % a.nco;  // output-channel blocks of one layer share a launch
  const int co0 = ib * UM_CB;
  const float* wimg = a.wimg + (int64_t)ib * a.img_stride;
  const int64_t HW = (int64_t)a.H * a.W, V = HW * a.D;

  // ---- one-time setup: barriers, weights, zero the ring tails, TMEM (allocated, then zeroed by the epilogue warps) --
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&plane_full[i], UM_NPROD); mbar_init(&plane_empty[i], 1); }
    for (int i = 0; i < UM_MT; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], UM_NEPI / UM_MT); }
    for (int i = 0; i < 4; ++i) mbar_init(&raw_full[i], 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < Cfg::W_BYTES / 16; i += UM_THREADS)
    reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(wimg) + i);
  // rows UM_PLANE .. UM_PFA of every chunk are only read by discarded accumulator rows; zero them once (no NaN patterns)
  for (int i = threadIdx.x; i < Cfg::NRING * 2 * Cfg::SCH * (UM_PFA - UM_PLANE); i += UM_THREADS) {
    const int t = i % (UM_PFA - UM_PLANE), c = i / (UM_PFA - UM_PLANE);
    reinterpret_cast<float4*>(ring)[c * UM_PFA + UM_PLANE + t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp == UM_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp < 8) {  // every MMA accumulates: the rotating blocks start from zero
    const uint32_t t0 = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * N);
#pragma unroll
    for (int c = 0; c < N; c += 16) tmem_st16_zero(t0 + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp > UM_MMA_WARP) {
    // =============================== producers ===============================
    // thread = one 16-byte K chunk (grp: 4 channels as tf32, 8 as fp16) x PPT fixed in-plane positions: everything but
    // the plane offset is loop invariant, and 56 loads per thread are in flight together (one or two memory latencies
    // per plane)
    constexpr int TPG = UM_NPROD / NCH;          // 24 (48) threads per channel chunk
    constexpr int PPT = UM_PLANE / TPG;          // 14 (7) positions per thread
    static_assert(TPG * NCH == UM_NPROD && PPT * TPG == UM_PLANE, "producer mapping must tile the plane exactly");
    const int tp = threadIdx.x - (UM_NEPI + 32);
    const int grp = tp / TPG, ti = tp - grp * TPG;
    const int cend = min(a.c0 + Cfg::KC, a.C1 + a.C2);
    if constexpr (Cfg::TMAIN) {
      // unit u = (plane u / NSUB, channels c0 + 16 * (u % NSUB) ..+15).  Thread 0 issues the TMA loads: two boxes of 8
      // channels per unit, each from the tensor that holds those channels (the host guarantees that no box straddles
      // the concatenation); everything outside the volume or beyond the channel count arrives as zeros.
      constexpr int NSUB = Cfg::NSUB, NRAW = Cfg::NRAW, NRING = Cfg::NRING;
      const int nunits = nsteps * NSUB;
      auto issue = [&](int u) {
        const int zi = z0 - 1 + u / NSUB, cu = a.c0 + 16 * (u % NSUB);
        uint64_t* bar = &raw_full[u % NRAW];
        float* dst = reinterpret_cast<float*>(rawbuf + (u % NRAW) * Cfg::RAW_SLOT_BYTES);
        mbar_expect_tx(bar, (uint32_t)Cfg::RAW_SLOT_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = cu + 8 * j;
          if (c < a.C1 || a.C2 == 0) tma_load_5d(dst + j * (Cfg::RAW_SLOT_BYTES / 8), &map1, bar, X0 - 4, Y0 - 1, zi, c, n);
          else tma_load_5d(dst + j * (Cfg::RAW_SLOT_BYTES / 8), &map2, bar, X0 - 4, Y0 - 1, zi, c - a.C1, n);
        }
      };
      if (tp == 0) {
        tma_prefetch_desc(&map1);
        if (a.C2) tma_prefetch_desc(&map2);
        for (int u = 0; u < NRAW && u < nunits; ++u) issue(u);
      }
      // thread = one 8-channel chunk of the unit (grp) x 7 in-plane positions
      constexpr int TPG = UM_NPROD / 2, PPT = UM_PLANE / TPG;
      static_assert(TPG * 2 == UM_NPROD && PPT * TPG == UM_PLANE, "producer mapping must tile the plane exactly");
      const int grp = tp / TPG, ti = tp - grp * TPG;
      uint32_t roff[PPT];   // byte offset of the thread's positions in a raw channel: row hy, column 3 + hx (the box starts at x0 - 4)
#pragma unroll
      for (int b = 0; b < PPT; ++b) {
        const int f = ti + TPG * b;
        const int hy = f / UM_PX, hx = f - hy * UM_PX;
        roff[b] = (uint32_t)(hy * UM_RAWX + hx + 3) * 4u;
      }
      constexpr uint32_t CHB = (UM_TY + 2) * UM_RAWX * 4;   // bytes of one raw channel
      const float sc = pow2f(scale_exp_from_amax(__ldg(a.amax_x)));
      const uint32_t raw_s = smem_u32(rawbuf) + (uint32_t)grp * 8u * CHB;
      const uint32_t ring_s = smem_u32(ring) + (uint32_t)(grp * UM_PFA + ti) * 16u;
      for (int u = 0; u < nunits; ++u) {
        const int rs = u % NRAW, slot = u % NRING;
        mbar_wait(&raw_full[rs], (u / NRAW) & 1);
        const uint32_t src = raw_s + (uint32_t)rs * Cfg::RAW_SLOT_BYTES;
        float v[PPT][8];
#pragma unroll
        for (int b = 0; b < PPT; ++b)
#pragma unroll
          for (int e = 0; e < 8; ++e) v[b][e] = lds_f32(src + roff[b] + (uint32_t)e * CHB);
        if (u >= NRING) mbar_wait(&plane_empty[slot], ((u / NRING) - 1) & 1);
        const uint32_t dst = ring_s + (uint32_t)slot * Cfg::SLOT_BYTES;
#pragma unroll
        for (int b = 0; b < PPT; ++b) {
          uint4 h, l;
          split_f16x2(v[b][0] * sc, v[b][1] * sc, h.x, l.x);
          split_f16x2(v[b][2] * sc, v[b][3] * sc, h.y, l.y);
          split_f16x2(v[b][4] * sc, v[b][5] * sc, h.z, l.z);
          split_f16x2(v[b][6] * sc, v[b][7] * sc, h.w, l.w);
          sts_v4(dst + (uint32_t)(TPG * b) * 16u, h);
          sts_v4(dst + (uint32_t)(TPG * b + Cfg::SCH * UM_PFA) * 16u, l);
        }
        fence_proxy_async();
        mbar_arrive(&plane_full[slot]);
        // The raw slot is handed back to the TMA unit only now: the stores above consumed every value loaded from it,
        // so no shared-memory read of this unit can still be queued when the asynchronous proxy overwrites the slot
        // (a barrier placed right after the loads let a rare late read lose that race: 1 of ~300 test runs).
        named_bar_sync(3, UM_NPROD);
        if (tp == 0 && u + NRAW < nunits) issue(u + NRAW);
      }
    } else if constexpr (!BF) {
      int oxy[PPT];
#pragma unroll
      for (int b = 0; b < PPT; ++b) {
        const int f = ti + TPG * b;
        const int hy = f / UM_PX, hx = f - hy * UM_PX;
        const int gy = Y0 - 1 + hy, gx = X0 - 1 + hx;
        oxy[b] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? gy * a.W + gx : -1;
      }
      const float* cb[4];
      bool cok[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = a.c0 + 4 * grp + e;
        cok[e] = c < cend;
        cb[e] = !cok[e] ? a.x1 : ((c < a.C1) ? a.x1 + ((int64_t)n * a.C1 + c) * V : a.x2 + ((int64_t)n * a.C2 + (c - a.C1)) * V);
      }
      for (int pi = 0; pi < nsteps; ++pi) {
        const int slot = pi % 3, use = pi / 3;
        if (use > 0) mbar_wait(&plane_empty[slot], (use - 1) & 1);
        const int zi = z0 - 1 + pi;
        float4* shi = reinterpret_cast<float4*>(ring + slot * Cfg::SLOT_BYTES) + grp * UM_PFA;
        float4* slo = shi + NCH * UM_PFA;
        const bool zok = zi >= 0 && zi < a.D;
        const int64_t zoff = (int64_t)zi * HW;
        float v[PPT][4];
#pragma unroll
        for (int b = 0; b < PPT; ++b)
#pragma unroll
          for (int e = 0; e < 4; ++e) v[b][e] = (zok && oxy[b] >= 0 && cok[e]) ? __ldg(cb[e] + zoff + oxy[b]) : 0.f;
#pragma unroll
        for (int b = 0; b < PPT; ++b) {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h[e] = __uint_as_float(__float_as_uint(v[b][e]) & 0xffffe000u);
            l[e] = v[b][e] - h[e];
          }
          shi[ti + TPG * b] = make_float4(h[0], h[1], h[2], h[3]);
          slo[ti + TPG * b] = make_float4(l[0], l[1], l[2], l[3]);
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        mbar_arrive(&plane_full[slot]);
      }
    } else {
      // eight channels per chunk: channel e of the chunk lives at b1 + e*V while e < esplit (first source) and at
      // b2 + e*V beyond (second source of a concatenation; b2 is biased by -esplit*V), e >= nv is zero padding
      const int cg0 = a.c0 + 8 * grp;
      const int nv = max(0, min(8, cend - cg0));
      const float* b1; const float* b2; int esplit;
      if (cg0 >= a.C1) {
        b1 = a.x2 ? a.x2 + ((int64_t)n * a.C2 + (cg0 - a.C1)) * V : a.x1; esplit = 8; b2 = b1;
      } else {
        b1 = a.x1 + ((int64_t)n * a.C1 + cg0) * V; esplit = min(8, a.C1 - cg0);
        b2 = a.x2 ? a.x2 + (int64_t)n * a.C2 * V - (int64_t)esplit * V : b1;
      }
      constexpr int BATCH = 7, NBATCH = PPT / BATCH;   // 56 loads in flight per thread
      static_assert(NBATCH * BATCH == PPT, "load batches must tile the thread's positions");
      const int Vi = (int)V;   // 8 * V < 2^31 (checked by the host): 32-bit element offsets, no per-channel pointers
      const float sc = pow2f(scale_exp_from_amax(__ldg(a.amax_x)));
      for (int pi = 0; pi < nsteps; ++pi) {
        const int slot = pi % 3, use = pi / 3;
        if (use > 0) mbar_wait(&plane_empty[slot], (use - 1) & 1);
        const int zi = z0 - 1 + pi;
        uint4* shi = reinterpret_cast<uint4*>(ring + slot * Cfg::SLOT_BYTES) + grp * UM_PFA;
        uint4* slo = shi + NCH * UM_PFA;
        const bool zok = zi >= 0 && zi < a.D;
        const float* p1 = b1 + (int64_t)zi * HW;
        const float* p2 = b2 + (int64_t)zi * HW;
        // the in-plane offsets are recomputed per plane (14 registers would otherwise be pinned for the whole walk and
        // the kernel spills); the volatile move keeps the compiler from hoisting them back out of the loop
        int tiv;
        asm volatile("mov.u32 %0, %1;" : "=r"(tiv) : "r"(ti));
#pragma unroll
        for (int q = 0; q < NBATCH; ++q) {
          float v[BATCH][8];
#pragma unroll
          for (int b = 0; b < BATCH; ++b) {
            const int f = tiv + TPG * (q * BATCH + b);
            const int hy = f / UM_PX, hx = f - hy * UM_PX;
            const int gy = Y0 - 1 + hy, gx = X0 - 1 + hx;
            const bool ok = zok && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
            const int o = gy * a.W + gx;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[b][e] = (ok && e < nv) ? __ldg((e < esplit ? p1 : p2) + (e * Vi + o)) : 0.f;
          }
#pragma unroll
          for (int b = 0; b < BATCH; ++b) {
            uint4 h, l;
            split_f16x2(v[b][0] * sc, v[b][1] * sc, h.x, l.x);
            split_f16x2(v[b][2] * sc, v[b][3] * sc, h.y, l.y);
            split_f16x2(v[b][4] * sc, v[b][5] * sc, h.z, l.z);
            split_f16x2(v[b][6] * sc, v[b][7] * sc, h.w, l.w);
            shi[ti + TPG * (q * BATCH + b)] = h;
            slo[ti + TPG * (q * BATCH + b)] = l;
          }
        }
        fence_proxy_async();
        mbar_arrive(&plane_full[slot]);
      }
    }
  } else if (warp == UM_MMA_WARP) {
    // =============================== MMA issue ===============================
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const uint32_t idesc = umma_idesc<BF>(128, N);
    const uint32_t ring_s = smem_u32(ring), sw_s = smem_u32(sw);
    constexpr uint32_t A_LBO = UM_PFA * 16, B_LBO = UM_WROWS * 16;
    const uint64_t adesc0 = umma_desc(0, A_LBO, 128), bdesc0 = umma_desc(0, B_LBO, 128);
    long long t_acc = 0, t_plane = 0, t_issue = 0;
    const long long t_begin = DBG ? clock64() : 0;
    if constexpr (Cfg::TMAIN) {
      // unit by unit: the MMAs of unit u read ring slot u % NRING and weight chunks 2 * (u % NSUB), +1; the accumulator
      // blocks rotate once per plane, so only the first unit of a plane waits for the epilogue and only the last one
      // hands the finished block over
      constexpr int NSUB = Cfg::NSUB, NRING = Cfg::NRING;
      for (int pi = 0; pi < nsteps; ++pi) {
        const uint32_t brot = sw_s + (uint32_t)(2 - pi % 3) * (NB * 16);
        const long long t1 = DBG ? clock64() : 0;
#pragma unroll 1
        for (int sub = 0; sub < NSUB; ++sub) {
          const int u = pi * NSUB + sub, slot = u % NRING;
          const long long t0 = DBG ? clock64() : 0;
          mbar_wait(&plane_full[slot], (u / NRING) & 1);
          if (DBG) t_plane += clock64() - t0;
          const uint32_t slot_s = ring_s + (uint32_t)slot * Cfg::SLOT_BYTES;
#pragma unroll 1
          for (int mt = 0; mt < UM_MT; ++mt) {
            const long long t2 = DBG ? clock64() : 0;
            if (sub == 0 && pi > 0) mbar_wait(&acc_empty[mt], (pi - 1) & 1);
            if (DBG) t_acc += clock64() - t2;
            tc_fence_after();
            if (elected) {
              const uint32_t dcol = tmem + (uint32_t)mt * N;
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const uint32_t arow = slot_s + (uint32_t)(mt * UM_MSTEP + ky * UM_PX) * 16;
                const uint32_t wt = brot + (uint32_t)(ky * NCH * UM_WROWS) * 16 + (uint32_t)(2 * sub) * B_LBO;
                const uint64_t a_hi = umma_desc_at(adesc0, arow);
                const uint64_t a_lo = umma_desc_at(adesc0, arow + Cfg::SCH * UM_PFA * 16);
                const uint64_t b_hi = umma_desc_at(bdesc0, wt);
                const uint64_t b_lo = umma_desc_at(bdesc0, wt + 3 * NCH * UM_WROWS * 16);
                umma_ss<BF>(dcol, a_hi, b_hi, idesc, 1u);
                if (!a.single_pass) {
                  umma_ss<BF>(dcol, a_lo, b_hi, idesc, 1u);
                  umma_ss<BF>(dcol, a_hi, b_lo, idesc, 1u);
                }
              }
              if (sub == NSUB - 1) umma_commit(&acc_full[mt]);
            }
            __syncwarp();
          }
          if (elected) umma_commit(&plane_empty[slot]);
          __syncwarp();
        }
        if (DBG) t_issue += clock64() - t1;
      }
    } else
    for (int pi = 0; pi < nsteps; ++pi) {
      const long long t0 = DBG ? clock64() : 0;
      mbar_wait(&plane_full[pi % 3], (pi / 3) & 1);
      const long long t1 = DBG ? clock64() : 0;
      t_plane += t1 - t0;
      const uint32_t slot_s = ring_s + (uint32_t)(pi % 3) * Cfg::SLOT_BYTES;
      const uint32_t brot = sw_s + (uint32_t)(2 - pi % 3) * (NB * 16);  // block b of this step holds kz = (pi - b) mod 3
#pragma unroll 1
      for (int mt = 0; mt < UM_MT; ++mt) {
        const long long t2 = DBG ? clock64() : 0;
        if (pi > 0) mbar_wait(&acc_empty[mt], (pi - 1) & 1);  // the epilogue has read + zeroed the block finished by step pi-1
        if (DBG) t_acc += clock64() - t2;
        tc_fence_after();
        if (elected) {
          const uint32_t dcol = tmem + (uint32_t)mt * N;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint32_t arow = slot_s + (uint32_t)(mt * UM_MSTEP + ky * UM_PX) * 16;  // (PX + mt*MSTEP) + (ky-1)*PX
            const uint32_t wt = brot + (uint32_t)(ky * NCH * UM_WROWS) * 16;
#pragma unroll
            for (int j2 = 0; j2 < NCH / 2; ++j2) {   // one instruction contracts two K chunks (8 tf32 / 16 fp16 channels)
              const uint64_t a_hi = umma_desc_at(adesc0, arow + (uint32_t)(2 * j2) * A_LBO);
              const uint64_t a_lo = umma_desc_at(adesc0, arow + (uint32_t)(2 * j2) * A_LBO + NCH * UM_PFA * 16);
              const uint64_t b_hi = umma_desc_at(bdesc0, wt + (uint32_t)(2 * j2) * B_LBO);
              const uint64_t b_lo = umma_desc_at(bdesc0, wt + (uint32_t)(2 * j2) * B_LBO + 3 * NCH * UM_WROWS * 16);
              umma_ss<BF>(dcol, a_hi, b_hi, idesc, 1u);
              if (!a.single_pass) {
                umma_ss<BF>(dcol, a_lo, b_hi, idesc, 1u);
                umma_ss<BF>(dcol, a_hi, b_lo, idesc, 1u);
              }
            }
          }
          umma_commit(&acc_full[mt]);
          if (mt == UM_MT - 1) umma_commit(&plane_empty[pi % 3]);
        }
        __syncwarp();
      }
      if (DBG) t_issue += clock64() - t1;
    }
    if (DBG && a.dbg && lane == 0) {
      atomicAdd(a.dbg + 0, (unsigned long long)t_acc); atomicAdd(a.dbg + 1, (unsigned long long)t_plane);
      atomicAdd(a.dbg + 2, (unsigned long long)t_issue); atomicAdd(a.dbg + 3, (unsigned long long)(clock64() - t_begin));
      atomicAdd(a.dbg + 4, (unsigned long long)nsteps); atomicAdd(a.dbg + 5, 1ull);
    }
  } else {
    // =============================== epilogue ===============================
    static_assert(UM_MT == 2 && UM_NEPI == 512 && UM_CB == 16, "one epilogue warp per (M tile, lane quadrant, channel half)");
    // two warps share every 32-row group and split its 16 output channels: the epilogue is latency bound per warp
    // (dependent TMEM load -> shuffle -> store chains), so twice the warps halve its share of the plane step
    // M tile mt covers plane rows [PX + mt*MSTEP, +128) and produces the outputs of its rows 1..126: the kx fold never
    // crosses an M tile, so the two tiles' epilogues run independently (a CTA-wide barrier here cost 1k of 4.7k cycles)
    const int w = warp & 3, mt = (warp >> 2) & 1, wg = warp & 7, ch0 = (warp >> 3) * 8;
    const int lrow = w * 32 + lane;                      // row within the M tile
    const int fc = UM_PX + mt * UM_MSTEP + lrow;
    const int hy = fc / UM_PX, hx = fc - hy * UM_PX;
    const int gy = Y0 - 1 + hy, gx = X0 - 1 + hx;
    const bool valid = lrow >= 1 && lrow <= UM_MSTEP && hx >= 1 && hx <= UM_TX && hy <= UM_TY && gy < a.H && gx < a.W;
    const uint32_t trow = tmem + ((uint32_t)(w * 32) << 16) + (uint32_t)mt * N;
    // bias in registers: a load inside the store loop cannot be hoisted past the stores (possible aliasing) and costs a
    // full L2 round trip per output channel (measured: 6.2k of the 7.1k cycles of a plane step)
    float bv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bv[c] = (a.last && a.bias && co0 + ch0 + c < a.Cout) ? __ldg(a.bias + co0 + ch0 + c) : 0.f;
    // fused activation as max(r, r * slope): leaky-ReLU / ReLU for 0 <= slope <= 1 (the host checks), identity at slope 1
    const float slope_eff = (a.last && a.act) ? a.slope : 1.f;
    const int nvc = max(0, min(8, a.Cout - co0 - ch0));   // valid output channels of this warp's half block
    // modes 1, 2: the operands were scaled by 2^kx and 2^kw; one exact multiplication undoes it (a combined exponent
    // below -126 means results under 2^-90: flushed to zero)
    float us = 1.f;
    if constexpr (BF) us = pow2f(max(-126, -(scale_exp_from_amax(__ldg(a.amax_x)) + scale_exp_from_amax(__ldg(a.amax_w)))));
    // The kx fold needs row f-1 (tap 0) and f+1 (tap 2).  Inside a warp they come by shuffle; the two rows at the warp's
    // ends need the neighbouring warp's edge values, which travel through shared memory.  A store -> barrier -> load
    // chain inside the step cost 1k of its 4.7k cycles (shared memory is saturated by the MMA operand fetch, every round
    // trip is several hundred cycles), so the two edge lanes DEFER their rows by one step: they keep the partial sum in
    // `pend`, and finish it at the next step, when the neighbour's values have long been published.
    const bool edge_lo = lane == 0 && w > 0, edge_hi = lane == 31 && w < 3;   // rows that wait for a neighbouring warp
    const bool edge_lane = lane == 0 || lane == 31;
    float pend[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) pend[c] = 0.f;
    long long e_wait = 0, e_tmem = 0, e_bar = 0, e_out = 0;
    const long long e_begin = DBG ? clock64() : 0;
    // 32-bit element offsets into the output (N * Cout * V < 2^32, checked by the host): one IMAD + one 64-bit LEA per
    // store instead of a 64-bit multiply-add chain.  Lanes outside the volume wrap around harmlessly: they never store.
    const uint32_t Vu = (uint32_t)V, HWu = (uint32_t)HW;
    const uint32_t obase = (uint32_t)(((int64_t)n * a.Cout + co0 + ch0) * V + (int64_t)gy * a.W + gx);
    const bool st_main = valid && !edge_lane, st_edge = valid && (edge_lo || edge_hi);
    const int nb = edge_lo ? wg - 1 : wg + 1, side = edge_lo ? 0 : 1;   // indices, not pointers: keeps LDS addressing
    float e[8];
    // fetch the neighbouring warp's edge values of step `pstep` (issued early, consumed by store_edges after the TMEM phase)
    auto fetch_edges = [&](int pstep) {
      const long long b0 = DBG ? clock64() : 0;
      named_bar_sync(1 + mt, UM_NEPI / UM_MT);           // every warp of this M tile has published that step's edges
      if (DBG) e_bar += clock64() - b0;
      if (edge_lo || edge_hi) {
#pragma unroll
        for (int c = 0; c < 8; ++c) e[c] = edge_s[pstep & 1][nb][side][ch0 + c];
      }
    };
    auto store_edges = [&](uint32_t opp) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float r = fmaxf(pend[c], pend[c] * slope_eff);
        if (st_edge && c < nvc) a.out[opp + (uint32_t)c * Vu] = r;
      }
    };
    for (int pi = 0; pi < nsteps; ++pi) {
      const int ol = pi - 2;                 // output plane completed by this step (local index), if >= 0
      const int blk = (pi + 1) % 3;          // = (pi - 2) mod 3
      const bool live = ol >= 0;             // ol < zcount always (nsteps = zcount + 2)
      const uint32_t op = obase + (uint32_t)(z0 + ol) * HWu;
      if (ol > 0) fetch_edges(pi - 1);       // the previous plane's edge rows: their neighbours' values are a step old
      // the previous chunks' partial output does not depend on this step's MMAs: fetch it (+ the bias) before waiting
      float old[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) old[c] = ((ACC && live && valid && c < nvc) ? __ldcg(a.out + (op + (uint32_t)c * Vu)) : 0.f) + bv[c];
      const long long e0 = DBG ? clock64() : 0;
      mbar_wait(&acc_full[mt], pi & 1);
      const long long e1 = DBG ? clock64() : 0;
      tc_fence_after();
      if (ol > 0 && (edge_lo || edge_hi)) {   // the edge values had the whole wait to arrive; folding them in here frees e[]
#pragma unroll
        for (int c = 0; c < 8; ++c) pend[c] += e[c];
      }
      float v[3][8];
      if (live) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) tmem_ld8(trow + (uint32_t)(blk * NB + kx * UM_CB + ch0), v[kx]);
        tmem_ld_wait();
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) tmem_st8_zero(trow + (uint32_t)(blk * NB + kx * UM_CB + ch0));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[mt]);  // block read and cleared: the next step may accumulate into it
      if (DBG) { e_wait += e1 - e0; e_tmem += clock64() - e1; }
      if (!live) continue;          // uniform over the CTA
      const long long e3 = DBG ? clock64() : 0;
      if (ol > 0) store_edges(op - HWu);
      if (lane == 31) {
#pragma unroll
        for (int c = 0; c < 8; ++c) edge_s[pi & 1][wg][0][ch0 + c] = BF ? v[0][c] * us : v[0][c];
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) edge_s[pi & 1][wg][1][ch0 + c] = BF ? v[2][c] * us : v[2][c];
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float left = __shfl_sync(0xffffffffu, v[0][c], (lane + 31) & 31);
        float right = __shfl_sync(0xffffffffu, v[2][c], (lane + 1) & 31);
        if (lane == 0) left = 0.f;      // wrapped around: the true neighbour lives in another warp (added by store_edges)
        if (lane == 31) right = 0.f;
        float r = v[1][c] + left + right;
        if constexpr (BF) r = fmaf(r, us, old[c]);
        else r += old[c];
        pend[c] = r;
        r = fmaxf(r, r * slope_eff);
        if (st_main && c < nvc) a.out[op + (uint32_t)c * Vu] = r;
      }
      if (DBG) e_out += clock64() - e3;
    }
    fetch_edges(nsteps - 1);
    if (edge_lo || edge_hi) {
#pragma unroll
      for (int c = 0; c < 8; ++c) pend[c] += e[c];
    }
    store_edges(obase + (uint32_t)(z0 + nsteps - 3) * HWu);
    if (DBG && a.dbg && threadIdx.x == 0) {
      atomicAdd(a.dbg + 9, (unsigned long long)e_bar); atomicAdd(a.dbg + 10, (unsigned long long)e_out);
      atomicAdd(a.dbg + 6, (unsigned long long)e_wait); atomicAdd(a.dbg + 7, (unsigned long long)e_tmem);
      atomicAdd(a.dbg + 8, (unsigned long long)(clock64() - e_begin));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == UM_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(Cfg::TMEM_COLS) : "memory");
}

// Weight image of one (16-output-channel block, 16-input-channel chunk):
//   [hi|lo][ky][ci/4][row = t*48 + kx*16 + co][4 floats],  t = 0..4 <-> kz = 2,1,0,2,1.
// Source indexing as repack_weights_kernel (a = this conv's input channel, b = its output channel).
__global__ void umma_prep_weights_kernel(const float* __restrict__ src, float* __restrict__ dst, int d1, int a_is_dim0, int flip,
                                         int A, int kc, int B, int b_off, int cb) {
  // grid (blocks, channel chunks, output-channel blocks): all images of a layer in one launch, image (ib, ik) at
  // dst + (ib * nk + ik) * W_FLOATS
  const int c0 = blockIdx.y * kc, co0 = blockIdx.z * cb;
  dst += (int64_t)(blockIdx.z * gridDim.y + blockIdx.y) * (UMMA_IMG_STRIDE_BYTES / 4);
  constexpr int UM_NCH = UmmaCfg<0>::NCH, UM_KC = UmmaCfg<0>::KC;
  const int total = UmmaCfg<0>::W_BYTES / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 3;
    int r = i >> 2;
    const int row = r % UM_WROWS; r /= UM_WROWS;
    const int j = r % UM_NCH; r /= UM_NCH;
    const int ky = r % 3;
    const int s = r / 3;
    const int t = row / UM_NB, kx = (row % UM_NB) / UM_CB, b = co0 + row % UM_CB, ai = c0 + 4 * j + e;
    const int kz = (t == 0 || t == 3) ? 2 : ((t == 1 || t == 4) ? 1 : 0);
    float v = 0.f;
    if (ai < A && ai < c0 + UM_KC && b < B) {
      const int tap = (kz * 3 + ky) * 3 + kx;
      const int ts = flip ? (26 - tap) : tap;
      const int i0 = a_is_dim0 ? ai : (b + b_off);
      const int i1 = a_is_dim0 ? (b + b_off) : ai;
      v = src[((int64_t)i0 * d1 + i1) * 27 + ts];
    }
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    dst[i] = s ? (v - hi) : hi;
  }
}

// Scaled fp16 hi/lo weight image of one (16-output-channel block, 32-input-channel chunk; the last chunk of a layer may
// be a 16-channel one, last_nch = 2):  [hi|lo][ky][ci/8][row = t*48 + kx*16 + co][8 halves], image (ib, ik) at
// dst + (ib * nk + ik) * UMMA_IMG_STRIDE_BYTES.  Chunk ik starts at input channel 32 * ik.  amax[1] = max|weight|.
__global__ void umma_prep_weights16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int d1, int a_is_dim0, int flip,
                                           int A, int B, int b_off, int last_nch, const float* __restrict__ amax_w) {
  const int ik = blockIdx.y, c0 = ik * 32, co0 = blockIdx.z * UM_CB;
  const int nch = (ik == (int)gridDim.y - 1) ? last_nch : 4;
  dst += (int64_t)(blockIdx.z * gridDim.y + ik) * (UMMA_IMG_STRIDE_BYTES / 2);
  const int half = 3 * nch * UM_WROWS * 8;   // elements of the hi (or lo) part
  const float sc = pow2f(scale_exp_from_amax(__ldg(amax_w)));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < half; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int r = i >> 3;
    const int row = r % UM_WROWS; r /= UM_WROWS;
    const int j = r % nch;
    const int ky = r / nch;
    const int t = row / UM_NB, kx = (row % UM_NB) / UM_CB, b = co0 + row % UM_CB, ai = c0 + 8 * j + e;
    const int kz = (t == 0 || t == 3) ? 2 : ((t == 1 || t == 4) ? 1 : 0);
    float v = 0.f;
    if (ai < A && b < B) {
      const int tap = (kz * 3 + ky) * 3 + kx;
      const int ts = flip ? (26 - tap) : tap;
      const int i0 = a_is_dim0 ? ai : (b + b_off);
      const int i1 = a_is_dim0 ? (b + b_off) : ai;
      v = src[((int64_t)i0 * d1 + i1) * 27 + ts] * sc;
    }
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    dst[i] = *reinterpret_cast<const uint16_t*>(&hi);
    dst[half + i] = *reinterpret_cast<const uint16_t*>(&lo);
  }
}
