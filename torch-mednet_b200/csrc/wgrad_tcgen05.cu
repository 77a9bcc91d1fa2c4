// Weight gradient of the 3x3x3 convolution on the 5th-generation tensor cores.
// ref: the autograd backward of nn.Conv3d at midasmednet/unet/components.py:8-9:
//      dW[co][ci][kd][kh][kw] = sum_v dY[v][co] * X[v + (kd-1, kh-1, kw-1)][ci]   (zero outside the volume)
//
// GEMM view.  The reduction (K) runs over VOXELS, so with NDHWC activations both operands are "MN-major"
// (channels contiguous, K strided): a TMA box of [voxels][channels] lands in shared memory exactly in the
// canonical MN-major swizzled layout (one voxel = one row), verified on the device by csrc/umma_lab.cu.
//   U  "unshifted" operand (A, M = 128 channels = 2 atoms of 64): a brick of TD x 16 x 8 voxels of the tensor
//      with MORE channels (dY or X).  Channels past the tensor end are zero-filled by TMA.
//   S  "shifted" operand (B): the (TD+2) x 18 x 10 HALO of the same brick of the OTHER tensor, 32 channels per
//      CTA, staged once.  The 27 taps are 27 shifted windows of that halo (descriptor start address only).
//      The three kw taps of one (kd, kh) are three windows one voxel row apart: the descriptor's
//      leading-dimension stride chains them into ONE instruction with N = 3 x 32 = 96.
//   D  one 128 x 96 fp32 accumulator per (kd, kh) group in TMEM; a CTA owns 5 (or 4) of the 9 groups
//      ("role"), 480 of the 512 TMEM columns, and accumulates over all the bricks it is given.
// Which tensor is U: the one with more channels.  With U = X, S = dY the sum is taken over X voxels u and
// the window offsets are mirrored:  dW[co][ci][k] = sum_u X[u][ci] * dY[u - (k-1)][co].
// Work decomposition: (U tile of 128 ch) x (S chunk of 32 ch) x (2 roles) x (split over bricks); every CTA
// writes one fp32 partial [128][480] and a second kernel sums the splits in a fixed order (deterministic)
// into the PyTorch (Cout, Cin, 3, 3, 3) layout.
// Cost model (measured): an M=128, N=96, K=16 MMA costs max(N/2, (4 KB + 32 N)/128) = 56 clk for
// 48 clk of math -> 86 % tensor-pipe ceiling; staging is (TD*32 + (TD+2)*11.25) KB per TD*16*5 MMAs.
#include <mutex>

#include "common.cuh"
#include "conv_impl.h"
#include "tc_common.cuh"

namespace mednet {

namespace {

int g_pair_planes = 1;        // mednet_tcgen05_set_option("wgrad_pair_planes", 0|1)
int g_d_fastest = 1;          // mednet_tcgen05_set_option("wgrad_d_fastest", 0|1)
int g_wt_fastest = 1;         // mednet_tcgen05_set_option("wgrad_wt_fastest", 0|1)
int g_dual = 1;               // mednet_tcgen05_set_option("wgrad_dual_issue", 0|1)
int g_class_merge = 1;        // mednet_tcgen05_set_option("wgrad_class_merge", 0|1): parity-class passes in one role, needed kw windows only
int g_reduce_s_fastest = 0;   // mednet_tcgen05_set_option("wgrad_reduce_s_fastest", 0|1): thread order of the split reduction (A/B switch)
int g_profile = 0;            // mednet_tcgen05_set_option("wgrad_profile", 0|1): wait-cycle counters, see wgrad_tc_kernel<PROF>

constexpr int WG_THREADS = 224;          // warp 0: TMA, warp 1: MMA issuer (+TMEM alloc), warps 2..5: epilogue, warp 6: second MMA issuer
constexpr int BR_H = 16, BR_W = 8;       // brick (h, w); depth TD
constexpr int HL_H = 18, HL_W = 10;      // halo plane
constexpr int CS = 32;                   // S channels per CTA (64-byte rows, SWIZZLE_64B)
constexpr int NCOLS = 3 * CS;            // N of one MMA (three chained kw taps)
constexpr int GROUPS0 = 5;               // (kd,kh) groups owned by role 0; role 1 owns the other 4
constexpr int PART_COLS = GROUPS0 * NCOLS;   // 480 columns per partial row
constexpr int STAGES = 2;

struct WgArgs {
  int N, D, H, W;
  int CU, CSn;                 // channels of U and S
  int TD;
  int tiles_d, tiles_h, tiles_w;
  int64_t bricks;              // N * tiles_d * tiles_h * tiles_w
  int u_tiles, s_chunks, ksplit;
  // transposed convolution (gather CONVT_B): the operand on the 2x grid (dY) is read through a stride-2 TMA map at
  // parity (cls >> 2, cls >> 1 & 1, cls & 1); one launch per parity class, only the (kd, kh) groups in gmask are needed
  int u_scale, s_scale, cls;
  uint32_t gmask;
  // Tap-group table of the parity-class passes (transposed conv, conv over an upsampled input): at most 4 of the 9
  // (kd, kh) groups are needed, so ONE role holds them all (nroles = 1): slot s holds group tab[0][s] (-1 = unused).  A
  // staged brick (110 KB at TD = 2) then feeds 4 MMAs per K step; split 2 + 2 over two roles every CTA pulled 61 B/clk
  // through L2 for 2 MMAs per K step and the class passes ran at 59 % of the MMA rate (ncu launch list r02g: 350 us for
  // 207 us of MMAs at 128x64 channels).  use_tab = 0: the fixed 0..4 / 4..8 ranges of the two roles.
  // Only the kw windows a class needs are chained: ncols = 32 x their count, kw0 = the first one.
  int use_tab, nroles, ncols, kw0;
  signed char tab[2][5];
  // second MMA-issuing thread (warp 6).  One thread sustains one tcgen05.mma per ~55-64 clk, the N = 96 MMA needs 56 clk of
  // shared-memory operand reads: two issuers, each owning a disjoint range of the CTA's tap-group accumulators, move the
  // bound from the issue rate to the operand rate (same scheme as the conv kernel's dual issue).
  int dual;
  int ut_base;                 // first U tile of this launch (the paired tail tile is launched separately)
  int pair_ok;                 // 1: U tiles with <= 64 real channels use the paired-plane mode (see kernel)
  int d_fastest;               // brick order inside a CTA: 1 = d fastest (halo planes reused from L2)
  int wt_fastest;              // block index order: 1 = work type fastest (bricks shared through L2), 0 = split fastest
  int64_t bricks_per_split;
  float* partial;              // [ksplit][worktype][128][PART_COLS]
};

// MMAs of one staged brick for the tap-group slots [GB, GE): TD planes x 8 pairs of h rows x (GE - GB) groups, straight-line
// (per-MMA predicates cost the issuing thread ~8 uniform-register moves each, taken or not: the group range is a template
// parameter and the per-brick choice is made once, outside)
template <int GB, int GE>
__device__ __forceinline__ void wg_issue_brick(int TD, uint32_t a_st, uint32_t b_st, uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                               const uint32_t (&tmem_g)[5], const uint32_t (&goff)[5], uint32_t a_dz16,
                                               uint32_t b_dz16, uint32_t a_hp16, uint32_t b_hp16) {
  for (int dz = 0; dz < TD; ++dz) {
    uint32_t a_lo = a_st + (uint32_t)dz * a_dz16, b_lo = b_st + (uint32_t)dz * b_dz16;
#pragma unroll
    for (int hp = 0; hp < 8; ++hp) {
#pragma unroll
      for (int g = GB; g < GE; ++g) tc::umma_bf16_lohi(tmem_g[g], a_lo, a_hi, b_lo + goff[g], b_hi, idesc, 1u);
      a_lo += a_hp16;
      b_lo += b_hp16;
    }
  }
}

// PROF: per-CTA cycle counters prof_out[blockIdx.x * 4 + i]: 0 MMA issuer total, 1 its wait for a filled stage,
// 2 TMA producer wait for a free stage, 3 bricks processed
template <bool PROF>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_s,
                const __grid_constant__ CUtensorMap map_u2, const WgArgs p, long long* __restrict__ prof_out) {
  long long* prof = PROF ? prof_out + (size_t)blockIdx.x * 4 : nullptr;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int u_atom_bytes = p.TD * 128 * 128;                       // one 64-channel atom of the brick
  const int u_bytes = 2 * u_atom_bytes;
  const int s_bytes_raw = (p.TD + 2) * HL_H * HL_W * (2 * CS);
  const int s_bytes = (s_bytes_raw + 1023) & ~1023;
  const int stage_bytes = u_bytes + s_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* done = bars + 2 * STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item of this CTA.  The work type varies FASTEST with blockIdx: the CTAs that need the same bricks of U and S
  // (other S chunk, other U tile, other tap-group role) are launched together and walk the same brick range at the same
  // pace, so a brick is fetched from HBM once and re-read from L2 (ncu, 64x128@64^3: 3.6 GB of DRAM reads for 0.8 GB of
  // operands with the split index fastest).
  const int worktypes = p.u_tiles * p.s_chunks * p.nroles;
  int wt = p.wt_fastest ? blockIdx.x % worktypes : blockIdx.x / p.ksplit;
  const int ks = p.wt_fastest ? blockIdx.x / worktypes : blockIdx.x % p.ksplit;
  const int wt0 = wt;          // partial buffer layout [ks][work type] whatever the launch order
  const int role = p.nroles == 2 ? (wt & 1) : 0; wt /= p.nroles;
  const int sc = wt % p.s_chunks;
  const int ut = wt / p.s_chunks + p.ut_base;
  // PAIRED-PLANE mode for a U tile with <= 64 real channels (64-channel layers; the 64-channel tail of a 192-channel
  // U): instead of leaving accumulator rows 64..127 empty, the second 64-row atom of A points at the SAME channels one
  // d-plane further (leading-dimension offset = one 16 KB plane of a (TD + 1)-plane brick).  With the S window at
  // (gd, gh), rows 0..63 accumulate tap (gd, gh, .) and rows 64..127 tap (gd - 1, gh, .): S windows gd = 1 (role 0) and
  // gd = 2 (role 1) cover all three kd taps with 3 MMAs per K step and CTA instead of 5 / 4.  The second half sums over
  // planes d0 + 1 .. d0 + TD, so the brick grid gets one extra layer at d0 = -TD (everything else there is zero fill).
  const bool paired = p.pair_ok && (p.CU - ut * 128) <= 64;
  // Unpaired tiles: role 0 holds tap groups 0..4, role 1 groups 4..8 (5 accumulator slots each).  The SHARED group 4 is
  // computed by role 0 on even bricks and by role 1 on odd ones (the reduce pass adds the two partial sums), so both
  // roles issue 4.5 MMAs per K step on average and the CTAs that read the same bricks stay in lockstep -- with a fixed
  // 5 / 4 split the faster role ran ahead and the shared bricks had to be fetched from HBM again (ncu: 2.5x the operand
  // bytes at 192x64@128^3).
  int tab_n = 0;
  if (p.use_tab)
    for (int i = 0; i < GROUPS0; ++i) tab_n += p.tab[role][i] >= 0 ? 1 : 0;
  const int ngroups = p.use_tab ? tab_n : (paired ? 3 : GROUPS0);
  const int g0 = paired ? 0 : (role == 0 ? 0 : GROUPS0 - 1);
  const int shared_slot = (paired || p.use_tab) ? -1 : (role == 0 ? GROUPS0 - 1 : 0);
  if (p.use_tab ? tab_n == 0 : (!paired && ((p.gmask >> g0) & ((1u << ngroups) - 1u)) == 0u))
    return;                                                                  // this role owns no needed tap group (whole CTA)
  const int tiles_d = p.tiles_d + (paired ? 1 : 0);
  const int64_t bricks = (int64_t)p.N * tiles_d * p.tiles_h * p.tiles_w;
  const int64_t per_split = (bricks + p.ksplit - 1) / p.ksplit;
  const int64_t b_begin = (int64_t)ks * per_split;
  int64_t b_end = b_begin + per_split;
  if (b_end > bricks) b_end = bricks;

  if (threadIdx.x == 0) {
    const uint32_t nissue = p.dual ? 2u : 1u;        // every issuer commits to the barriers its MMAs release
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], nissue); }
    tc::mbar_init(done, nissue);
    tc::fence_barrier_init();
    tc::tma_prefetch_desc(&map_u);
    tc::tma_prefetch_desc(&map_s);
    tc::tma_prefetch_desc(&map_u2);
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512u);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  {
    // every accumulator slot starts at zero, so every MMA accumulates: no first-MMA special case in the issue loop, and the
    // shared slot (which this CTA's first brick may not touch) needs no separate treatment
    if (warp >= 2 && warp <= 5) {
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      for (int j = 0; j < ngroups * NCOLS; j += 16) tc::tmem_st_x16_zero(taddr + (uint32_t)j);
      tc::tmem_st_wait();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      uint32_t it = 0;
      long long w_empty = 0;
      for (int64_t b = b_begin; b < b_end; ++b, ++it) {
        const uint32_t st = it % STAGES, ph = (it / STAGES) & 1u;
        int64_t t = b;
        // d varies fastest: consecutive bricks of a CTA share TD of the TD + 2 halo planes of S while they are still in L2
        // (w fastest re-read them 128 bricks later, i.e. from HBM: 3.4x the operand bytes at 192x64@128^3)
        int d0, h0, w0, n;
        if (p.d_fastest) {
          d0 = (int)(t % tiles_d) * p.TD - (paired ? p.TD : 0); t /= tiles_d;
          w0 = (int)(t % p.tiles_w) * BR_W; t /= p.tiles_w;
          h0 = (int)(t % p.tiles_h) * BR_H;
          n = (int)(t / p.tiles_h);
        } else {
          w0 = (int)(t % p.tiles_w) * BR_W; t /= p.tiles_w;
          h0 = (int)(t % p.tiles_h) * BR_H; t /= p.tiles_h;
          d0 = (int)(t % tiles_d) * p.TD - (paired ? p.TD : 0);
          n = (int)(t / tiles_d);
        }
        if (PROF) { const long long t0 = clock64(); tc::mbar_wait(&empty[st], ph ^ 1u); w_empty += clock64() - t0; }
        else tc::mbar_wait(&empty[st], ph ^ 1u);
        uint8_t* dst = smem + (size_t)st * stage_bytes;
        if (paired) {
          tc::mbar_arrive_expect_tx(&full[st], (uint32_t)((p.TD + 1) * 128 * 128 + s_bytes_raw));
          tc::tma_load_5d(dst, &map_u2, &full[st], ut * 128, w0, h0, d0, n);
          tc::tma_load_5d(dst + u_bytes, &map_s, &full[st], sc * CS, w0 - 1, h0 - 1, d0 - 1, n);
          continue;
        }
        tc::mbar_arrive_expect_tx(&full[st], (uint32_t)(u_bytes + s_bytes_raw));
        const int us = p.u_scale, ss = p.s_scale;
        const int upd = us > 1 ? (p.cls >> 2) : 0, uph = us > 1 ? ((p.cls >> 1) & 1) : 0, upw = us > 1 ? (p.cls & 1) : 0;
        const int spd = ss > 1 ? (p.cls >> 2) : 0, sph = ss > 1 ? ((p.cls >> 1) & 1) : 0, spw = ss > 1 ? (p.cls & 1) : 0;
        tc::tma_load_5d(dst, &map_u, &full[st], ut * 128, us * w0 + upw, us * h0 + uph, us * d0 + upd, n);
        tc::tma_load_5d(dst + u_atom_bytes, &map_u, &full[st], ut * 128 + 64, us * w0 + upw, us * h0 + uph, us * d0 + upd, n);
        tc::tma_load_5d(dst + u_bytes, &map_s, &full[st], sc * CS, ss * (w0 - 1) + spw, ss * (h0 - 1) + sph,
                        ss * (d0 - 1) + spd, n);
      }
      if (PROF) prof[2] = w_empty;
    }
  } else if (warp == 1 || warp == 6) {
    // ===================== MMA issuer(s) =====================
    const int issuer = warp == 1 ? 0 : 1;
    if ((issuer == 0 || p.dual) && tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(128, p.ncols, 1, 1, 1, 1);
      // per-group window offset inside the halo, in descriptor units (16 bytes)
      uint32_t goff[GROUPS0];
#pragma unroll
      for (int g = 0; g < GROUPS0; ++g) {
        const int gg = p.use_tab ? (g < ngroups ? p.tab[role][g] : p.tab[role][0]) : g0 + (g < ngroups ? g : 0);
        const int gd = paired ? 1 + role : gg / 3, gh = paired ? gg : gg - gd * 3;
        goff[g] = (uint32_t)(((gd * HL_H + gh) * HL_W + p.kw0) * (2 * CS)) >> 4;
      }
      const uint32_t s_base = tc::smem_u32(smem);
      // A: MN-major, 128-byte rows, atoms u_atom_bytes apart (paired mode: the same channels one 16 KB d-plane further),
      //    8-row groups 1024 B apart.  B: MN-major, 64-byte rows, three chained windows one row (64 B) apart, 8-row
      //    groups = next halo row.
      const uint64_t da0 = tc::make_smem_desc(s_base, paired ? 128u * 128u : (uint32_t)u_atom_bytes, 1024u, 0, tc::SWZ_128B);
      const uint64_t db0 = tc::make_smem_desc(s_base + (uint32_t)u_bytes, (uint32_t)(2 * CS), (uint32_t)(HL_W * 2 * CS), 0,
                                              tc::SWZ_64B);
      const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
      const uint32_t a_lo0 = (uint32_t)da0, b_lo0 = (uint32_t)db0;
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t a_dz16 = (128u * 128u) >> 4, b_dz16 = (uint32_t)(HL_H * HL_W * 2 * CS) >> 4;
      const uint32_t a_hp16 = (16u * 128u) >> 4, b_hp16 = (uint32_t)(2 * HL_W * 2 * CS) >> 4;
      uint32_t tmem_g[GROUPS0];
      uint32_t base_mask = 0;
#pragma unroll
      for (int g = 0; g < GROUPS0; ++g) {
        tmem_g[g] = tmem_base + (uint32_t)(g * NCOLS);
        if (g < ngroups && (paired || p.use_tab || ((p.gmask >> (g0 + g)) & 1u))) base_mask |= 1u << g;
      }
      const uint32_t shared_bit = shared_slot >= 0 ? (1u << shared_slot) : 0u;
      uint32_t it = 0;
      long long w_full = 0;
      const long long t_begin = PROF ? clock64() : 0;
      for (int64_t b = b_begin; b < b_end; ++b, ++it) {
        const uint32_t st = it % STAGES, ph = (it / STAGES) & 1u;
        if (PROF) { const long long t0 = clock64(); tc::mbar_wait(&full[st], ph); w_full += clock64() - t0; }
        else tc::mbar_wait(&full[st], ph);
        tc::tc_fence_after();
        // The issuing thread is a serial instruction stream (~55 clk per tcgen05.mma at best): descriptors are built once
        // per kernel, per MMA only 32-bit address words are advanced, and which tap groups this brick feeds is one
        // bit mask evaluated per brick, not per MMA.
        const uint32_t a_st = a_lo0 + st * stage16, b_st = b_lo0 + st * stage16;
        const bool shared_mine = (int)(b & 1) == role;          // which role computes the shared group for this brick
#define WG_BRICK(GB, GE) wg_issue_brick<GB, GE>(p.TD, a_st, b_st, a_hi, b_hi, idesc, tmem_g, goff, a_dz16, b_dz16, a_hp16, b_hp16)
        if (p.dual) {
          // each issuer owns a disjoint range of the brick's tap groups (their accumulators are disjoint TMEM columns)
          if (paired) { if (issuer == 0) WG_BRICK(0, 2); else WG_BRICK(2, 3); }
          else if (p.use_tab && ngroups == 4) { if (issuer == 0) WG_BRICK(0, 2); else WG_BRICK(2, 4); }
          else if (p.use_tab && ngroups == 2) { if (issuer == 0) WG_BRICK(0, 1); else WG_BRICK(1, 2); }
          else if (p.use_tab && ngroups == 1) { if (issuer == 0) WG_BRICK(0, 1); }
          else if (base_mask == 0x1fu) {
            if (shared_mine) { if (issuer == 0) WG_BRICK(0, 3); else WG_BRICK(3, 5); }
            else if (role == 0) { if (issuer == 0) WG_BRICK(0, 2); else WG_BRICK(2, 4); }
            else { if (issuer == 0) WG_BRICK(1, 3); else WG_BRICK(3, 5); }
          } else {
            const uint32_t act = (shared_mine ? base_mask : (base_mask & ~shared_bit)) & (issuer == 0 ? 0x15u : 0x0au);
            for (int dz = 0; dz < p.TD; ++dz) {
              uint32_t a_lo = a_st + (uint32_t)dz * a_dz16, b_lo = b_st + (uint32_t)dz * b_dz16;
              for (int hp = 0; hp < 8; ++hp) {
#pragma unroll
                for (int g = 0; g < GROUPS0; ++g)
                  if ((act >> g) & 1u) tc::umma_bf16_lohi(tmem_g[g], a_lo, a_hi, b_lo + goff[g], b_hi, idesc, 1u);
                a_lo += a_hp16;
                b_lo += b_hp16;
              }
            }
          }
        } else if (paired) {
          WG_BRICK(0, 3);
        } else if (p.use_tab && ngroups == 4) {
          WG_BRICK(0, 4);
        } else if (p.use_tab && ngroups == 2) {
          WG_BRICK(0, 2);
        } else if (p.use_tab && ngroups == 1) {
          WG_BRICK(0, 1);
        } else if (base_mask == 0x1fu) {
          if (shared_mine) WG_BRICK(0, 5);
          else if (role == 0) WG_BRICK(0, 4);
          else WG_BRICK(1, 5);
        } else {
          // any other group subset (gmask with the fixed ranges): predicated, not used by the shipped plans
          const uint32_t act = shared_mine ? base_mask : (base_mask & ~shared_bit);
          for (int dz = 0; dz < p.TD; ++dz) {
            uint32_t a_lo = a_st + (uint32_t)dz * a_dz16, b_lo = b_st + (uint32_t)dz * b_dz16;
            for (int hp = 0; hp < 8; ++hp) {
#pragma unroll
              for (int g = 0; g < GROUPS0; ++g)
                if ((act >> g) & 1u) tc::umma_bf16_lohi(tmem_g[g], a_lo, a_hi, b_lo + goff[g], b_hi, idesc, 1u);
              a_lo += a_hp16;
              b_lo += b_hp16;
            }
          }
        }
        tc::umma_commit(&empty[st]);
      }
      tc::umma_commit(done);
#undef WG_BRICK
      if (PROF && issuer == 0) { tc::mbar_wait(done, 0); prof[0] = clock64() - t_begin; prof[1] = w_full; prof[3] = b_end - b_begin; }
    }
  } else {
    // ===================== epilogue (once per CTA) =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float* prow = p.partial + (((size_t)ks * worktypes + wt0) * 128 + m) * PART_COLS;
    if (b_end > b_begin) {
      tc::mbar_wait(done, 0);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int j = 0; j < ngroups * NCOLS; j += 16) {
        uint32_t r[16];
        tc::tmem_ld_x16(taddr + (uint32_t)j, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(prow + j + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                 __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
    } else {
      for (int j = 0; j < ngroups * NCOLS; j += 4) *reinterpret_cast<float4*>(prow + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512u);
}

struct WgTab { int use; signed char where[9]; };     // where[group] = role * 8 + slot

// dw[co][ci][kd][kh][kw] (+)= sum over splits of the partial accumulators, fixed order
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Cout, int Cin,
                                       int u_is_x, int s_chunks, int ksplit, int worktypes, int accumulate, int cls,
                                       int CU, int pair_ok, int seg1_tile, int ksplit1, int worktypes1, int64_t seg1_offset,
                                       int upconv, WgTab tab, int nroles, int kw0, int dw_ld, int dw_c0, int dw_transposed,
                                       int s_fastest) {
  // partial buffer: segment 0 = U tiles [0, seg1_tile) as [ksplit][worktypes][128][PART_COLS]; segment 1 (the paired
  // tail tile, if any) starts at seg1_offset floats with its own split factor
  const int64_t total = (int64_t)Cout * Cin * 27;
  // Thread order.  Default: the output order (tap fastest) -- the read-modify-write of dw is contiguous, every lane's load
  // of a split is its own 32-byte sector (the neighbouring sectors are used by the next warps, out of L2).  s_fastest: S
  // channel fastest, then tap, then U channel -- a warp reads one 128-byte line per split and scatters its results at a
  // 27-float stride.  Measured on one box (cfg-3, profiles/r02/r02s_bench_cfg3_reduce_order_*.json): 73.3 ms/step with
  // the output order against 74.9 with s_fastest, so the coalesced reads do not pay for the scattered read-modify-write;
  // kept as an A/B switch.  The 37 reduce launches cost 2.0 ms per step (profiles/r02/kernels_r02q.txt).
  const int CSn = u_is_x ? Cout : Cin;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    int tap, co, ci;
    if (s_fastest) {
      const int cs_f = (int)(j % CSn);
      tap = (int)((j / CSn) % 27);
      const int cu_f = (int)(j / (27 * (int64_t)CSn));
      co = u_is_x ? cs_f : cu_f;
      ci = u_is_x ? cu_f : cs_f;
    } else {                                                       // output order (tap fastest): the A/B baseline
      tap = (int)(j % 27);
      ci = (int)((j / 27) % Cin);
      co = (int)(j / (27 * (int64_t)Cin));
    }
    const int64_t i = ((int64_t)co * Cin + ci) * 27 + tap;         // index in the dense (Cout, Cin, 27) gradient
    int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    if (cls >= 0 && upconv) {
      // conv over the nearest-upsampled input: EVERY fine tap o receives a share from every dY parity class p, namely the
      // class partial at window tap 1 - c(p, o) with c(0, .) = (-1, 0, 0), c(1, .) = (0, 0, +1)  (see MEDNET_GATHER_UPCONV_*)
      const int pd = cls >> 2, ph = (cls >> 1) & 1, pw = cls & 1;
      kd = pd ? (kd == 2 ? 0 : 1) : (kd == 0 ? 2 : 1);
      kh = ph ? (kh == 2 ? 0 : 1) : (kh == 0 ? 2 : 1);
      kw = pw ? (kw == 2 ? 0 : 1) : (kw == 0 ? 2 : 1);
    } else if (cls >= 0) {
      // transposed conv: kernel tap k belongs to dY parity (k != 1) per axis and window tap (k == 0 ? 0 : 1)
      if ((((kd != 1) << 2) | ((kh != 1) << 1) | (kw != 1)) != cls) continue;
      kd = kd == 0 ? 0 : 1; kh = kh == 0 ? 0 : 1; kw = kw == 0 ? 0 : 1;
    }
    int cu, cs;
    if (u_is_x) { cu = ci; cs = co; kd = 2 - kd; kh = 2 - kh; kw = 2 - kw; }
    else { cu = co; cs = ci; }
    int role, gl, row = cu & 127;
    bool shared = false;
    if (pair_ok && (CU - (cu >> 7) * 128) <= 64) {
      // paired-plane tile: role 0 (S window gd = 1) holds tap kd = 1 in rows 0..63 and kd = 0 in rows 64..127,
      // role 1 (gd = 2) holds kd = 2 in rows 0..63; the (kd, kh) group index inside the CTA is kh
      role = kd == 2 ? 1 : 0;
      gl = kh;
      row = (cu & 63) + (kd == 0 ? 64 : 0);
    } else {
      const int g = kd * 3 + kh;               // role 0: groups 0..4 in slots 0..4; role 1: groups 4..8 in slots 0..4
      if (tab.use) {                           // balanced table: group g lives in (role, slot) = where[g]
        role = tab.where[g] >> 3;
        gl = tab.where[g] & 7;
      } else {
        role = g >= GROUPS0 ? 1 : 0;
        gl = g - role * (GROUPS0 - 1);
        shared = g == GROUPS0 - 1;             // group 4: role 0 slot 4 (even bricks) + role 1 slot 0 (odd bricks)
      }
    }
    const int tile = cu >> 7;
    const bool s1 = tile >= seg1_tile;
    const int wt = ((tile - (s1 ? seg1_tile : 0)) * s_chunks + cs / CS) * nroles + role;
    const int nk = s1 ? ksplit1 : ksplit, nw = s1 ? worktypes1 : worktypes;
    const float* src = partial + (s1 ? seg1_offset : 0) + ((size_t)wt * 128 + row) * PART_COLS + gl * NCOLS + (kw - kw0) * CS + (cs % CS);
    // four interleaved partial sums: the up-to-148 split partials are independent loads, a single running sum would
    // serialise their latencies (the reduce launches were latency-bound: 37 of them cost 1.5 ms per step); the order is
    // still fixed, hence deterministic
    const size_t kst = (size_t)nw * 128 * PART_COLS;
    auto sum_splits = [&](const float* s0) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
      for (; k + 3 < nk; k += 4) {
        a0 += s0[(size_t)k * kst];
        a1 += s0[(size_t)(k + 1) * kst];
        a2 += s0[(size_t)(k + 2) * kst];
        a3 += s0[(size_t)(k + 3) * kst];
      }
      for (; k < nk; ++k) a0 += s0[(size_t)k * kst];
      return (a0 + a1) + (a2 + a3);
    };
    float acc = sum_splits(src);
    if (shared)                                // the other role's share: work type + 1, slot 0
      acc += sum_splits(src + (size_t)128 * PART_COLS - (size_t)gl * NCOLS);
    // destination: dense (Ca, Cb, 27), or a channel range of a wider / transposed gradient tensor (mednet_wgrad_params)
    const int64_t o = dw_ld == 0 ? i
                                 : (dw_transposed ? ((int64_t)ci * dw_ld + dw_c0 + co) * 27 + tap
                                                  : ((int64_t)co * dw_ld + dw_c0 + ci) * 27 + tap);
    dw[o] = (accumulate || (upconv && cls > 0)) ? dw[o] + acc : acc;      // upconv: classes 1..7 add to what class 0 wrote
  }
}

struct WgSeg { int ut_base, u_tiles, ksplit, grid; size_t offset_floats; };
struct WgPlan {
  WgArgs a;
  int u_is_x;
  size_t partial_bytes, smem;
  int nseg;
  WgSeg seg[2];              // [0] full 128-channel U tiles, [1] the paired 64-channel tail tile (own split factor)
};

bool plan_wgrad(const mednet_wgrad_params* q, WgPlan* out) {
  if (q->dtype != MEDNET_BF16) return false;
  // a = x on the small grid, b = dY on the 2x grid (transposed conv, or the conv over a nearest-upsampled input)
  const bool convt = q->gather == MEDNET_GATHER_CONVT_B || q->gather == MEDNET_GATHER_UPCONV_B;
  if (q->gather != MEDNET_GATHER_CONV3 && !convt) return false;
  if (convt && (q->Db != 2 * q->Da || q->Hb != 2 * q->Ha || q->Wb != 2 * q->Wa)) return false;
  if (!mednet_device_has_tcgen05()) return false;
  if (((uintptr_t)q->a | (uintptr_t)q->b) & 15) return false;
  WgPlan pl;
  // U = the operand with more channels (ties: dY); S = the other one, in chunks of 32 channels
  pl.u_is_x = q->Cb > q->Ca ? 1 : 0;
  const int CU = pl.u_is_x ? q->Cb : q->Ca, CSn = pl.u_is_x ? q->Ca : q->Cb;
  // any multiple of 8 channels (16-byte global strides): TMA boxes reaching past the channel extent are zero-filled, so a
  // 32- or 96-channel U simply leaves rows of the 128-row accumulator empty and a 16-channel S leaves columns empty
  if (CU % 8 != 0 || CSn % 8 != 0) return false;
  WgArgs& a = pl.a;
  a.N = q->N; a.D = q->Da; a.H = q->Ha; a.W = q->Wa; a.CU = CU; a.CSn = CSn;
  a.u_scale = (convt && pl.u_is_x) ? 2 : 1;                   // "u_is_x": U is operand b
  a.s_scale = (convt && !pl.u_is_x) ? 2 : 1;
  a.cls = 0; a.gmask = 0x1ffu; a.use_tab = 0; a.dual = g_dual;
  a.nroles = (convt && g_class_merge) ? 1 : 2;                                   // parity-class passes: all (<= 4) tap groups in one role
  a.ncols = NCOLS; a.kw0 = 0;
  for (int r = 0; r < 2; ++r)
    for (int i = 0; i < 5; ++i) a.tab[r][i] = -1;
  a.pair_ok = (!convt && g_pair_planes) ? 1 : 0;
  a.TD = q->Da >= 2 ? 2 : 1;
  a.tiles_d = ceil_div(a.D, a.TD); a.tiles_h = ceil_div(a.H, BR_H); a.tiles_w = ceil_div(a.W, BR_W);
  a.bricks = (int64_t)a.N * a.tiles_d * a.tiles_h * a.tiles_w;
  a.s_chunks = ceil_div(CSn, CS);
  const int tiles_all = ceil_div(CU, 128);
  const bool tail_paired = a.pair_ok && (CU - (tiles_all - 1) * 128) <= 64;
  const int sms = sm_count_cached();
  pl.nseg = 0;
  size_t off = 0;
  auto add_seg = [&](int ut_base, int u_tiles, bool paired) {
    WgSeg& sg = pl.seg[pl.nseg++];
    sg.ut_base = ut_base; sg.u_tiles = u_tiles;
    const int worktypes = u_tiles * a.s_chunks * a.nroles;
    const int64_t bricks = (int64_t)a.N * (a.tiles_d + (paired ? 1 : 0)) * a.tiles_h * a.tiles_w;
    // one CTA is resident per SM (224 KB of shared memory): AT MOST two full waves, never a third, nearly empty one
    // (rounding the split factor up gave e.g. 312 CTAs = 148 + 148 + 16 for 384x128 channels: 30 % of the launch idle)
    int64_t ksplit = (2 * sms) / worktypes;
    if (ksplit > bricks) ksplit = bricks;
    if (ksplit < 1) ksplit = 1;
    sg.ksplit = (int)ceil_div64(bricks, ceil_div64(bricks, ksplit));
    sg.grid = worktypes * sg.ksplit;
    sg.offset_floats = off;
    off += (size_t)sg.grid * 128 * PART_COLS;
  };
  if (tail_paired) {
    if (tiles_all > 1) add_seg(0, tiles_all - 1, false);
    add_seg(tiles_all - 1, 1, true);
  } else {
    add_seg(0, tiles_all, false);
  }
  pl.partial_bytes = align_up(off * sizeof(float), 256);
  const size_t u_bytes = (size_t)2 * a.TD * 128 * 128;
  const size_t s_bytes = align_up((size_t)(a.TD + 2) * HL_H * HL_W * 2 * CS, 1024);
  pl.smem = 1024 + STAGES * (u_bytes + s_bytes) + 128;
  if (pl.smem > 227 * 1024) return false;
  *out = pl;
  return true;
}

}  // namespace

void tc_wgrad_set_wt_fastest(int v) { g_wt_fastest = v ? 1 : 0; }
void tc_wgrad_set_pair_planes(int v) { g_pair_planes = v ? 1 : 0; }
void tc_wgrad_set_d_fastest(int v) { g_d_fastest = v ? 1 : 0; }
void tc_wgrad_set_profile(int v) { g_profile = v ? 1 : 0; }
void tc_wgrad_set_dual(int v) { g_dual = v ? 1 : 0; }
void tc_wgrad_set_class_merge(int v) { g_class_merge = v ? 1 : 0; }
void tc_wgrad_set_reduce_s_fastest(int v) { g_reduce_s_fastest = v ? 1 : 0; }
size_t tc_wgrad_profile_offset(const mednet_wgrad_params* q) {
  WgPlan pl;
  if (!plan_wgrad(q, &pl)) return 0;
  return pl.partial_bytes + colsum_workspace_bytes(q);
}

bool tc_wgrad_supported(const mednet_wgrad_params* q) {
  WgPlan pl;
  return plan_wgrad(q, &pl);
}

size_t tc_wgrad_workspace_bytes(const mednet_wgrad_params* q) {
  WgPlan pl;
  if (!plan_wgrad(q, &pl)) return 0;
  return pl.partial_bytes + colsum_workspace_bytes(q) + (g_profile ? 2 * 4 * 1024 * sizeof(long long) : 0);
}

int tc_wgrad(const mednet_wgrad_params* q, void* workspace, cudaStream_t st) {
  WgPlan pl;
  if (!plan_wgrad(q, &pl)) return MEDNET_EUNSUPPORTED;
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return MEDNET_ENODRIVER;
  WgArgs& a = pl.a;
  a.wt_fastest = g_wt_fastest;
  a.d_fastest = g_d_fastest;
  const void* u_ptr = pl.u_is_x ? q->b : q->a;
  const void* s_ptr = pl.u_is_x ? q->a : q->b;
  CUtensorMap map_u, map_s;
  const bool upconv = q->gather == MEDNET_GATHER_UPCONV_B;
  const bool convt = q->gather == MEDNET_GATHER_CONVT_B || upconv;
  {
    const cuuint64_t C = (cuuint64_t)a.CU;
    const cuuint64_t GD = pl.u_is_x ? q->Db : q->Da, GH = pl.u_is_x ? q->Hb : q->Ha, GW = pl.u_is_x ? q->Wb : q->Wa;
    const cuuint32_t es = (cuuint32_t)a.u_scale;
    cuuint64_t dims[5] = {C, GW, GH, GD, (cuuint64_t)a.N};
    cuuint64_t strides[4] = {C * 2, GW * C * 2, GH * GW * C * 2, GD * GH * GW * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)BR_W * es, (cuuint32_t)BR_H * es, (cuuint32_t)a.TD * es, 1};
    cuuint32_t estr[5] = {1, es, es, es, 1};
    if (enc(&map_u, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(u_ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MEDNET_EUNSUPPORTED;
  }
  CUtensorMap map_u2 = map_u;
  if (a.pair_ok) {       // paired-plane bricks: one 64-channel atom, TD + 1 planes
    const cuuint64_t C = (cuuint64_t)a.CU;
    cuuint64_t dims[5] = {C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.N};
    cuuint64_t strides[4] = {C * 2, (cuuint64_t)a.W * C * 2, (cuuint64_t)a.H * a.W * C * 2, (cuuint64_t)a.D * a.H * a.W * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)BR_W, (cuuint32_t)BR_H, (cuuint32_t)(a.TD + 1), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(&map_u2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(u_ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MEDNET_EUNSUPPORTED;
  }
  {
    const cuuint64_t C = (cuuint64_t)a.CSn;
    const cuuint64_t GD = pl.u_is_x ? q->Da : q->Db, GH = pl.u_is_x ? q->Ha : q->Hb, GW = pl.u_is_x ? q->Wa : q->Wb;
    const cuuint32_t es = (cuuint32_t)a.s_scale;
    cuuint64_t dims[5] = {C, GW, GH, GD, (cuuint64_t)a.N};
    cuuint64_t strides[4] = {C * 2, GW * C * 2, GH * GW * C * 2, GD * GH * GW * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)CS, (cuuint32_t)HL_W * es, (cuuint32_t)HL_H * es, (cuuint32_t)(a.TD + 2) * es, 1};
    cuuint32_t estr[5] = {1, es, es, es, 1};
    if (enc(&map_s, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(s_ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MEDNET_EUNSUPPORTED;
  }
  static std::mutex mu;
  static size_t configured = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (pl.smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
      if (e != cudaSuccess) return (int)e;
      configured = pl.smem;
    }
  }
  const int64_t total = (int64_t)q->Ca * q->Cb * 27;
  const WgSeg& s0 = pl.seg[0];
  const WgSeg& s1 = pl.seg[pl.nseg - 1];
  // reduce-kernel view: tiles >= seg1_tile live in the second segment (none when there is a single segment)
  const int seg1_tile = pl.nseg == 2 ? s1.ut_base : (1 << 20);
  auto launch_all = [&]() -> int {
    for (int i = 0; i < pl.nseg; ++i) {
      const WgSeg& sg = pl.seg[i];
      a.ut_base = sg.ut_base; a.u_tiles = sg.u_tiles; a.ksplit = sg.ksplit;
      a.partial = (float*)workspace + sg.offset_floats;
      if (g_profile)     // counters of segment i behind the partial sums and the bias scratch (workspace sized for it)
        wgrad_tc_kernel<true><<<(unsigned)sg.grid, WG_THREADS, pl.smem, st>>>(
            map_u, map_s, map_u2, a, (long long*)((char*)workspace + pl.partial_bytes + colsum_workspace_bytes(q)) + (size_t)i * 4 * 1024);
      else
        wgrad_tc_kernel<false><<<(unsigned)sg.grid, WG_THREADS, pl.smem, st>>>(map_u, map_s, map_u2, a, nullptr);
      MEDNET_LAUNCH_CHECK();
    }
    return MEDNET_OK;
  };
  WgTab wtab;
  wtab.use = 0;
  for (int i = 0; i < 9; ++i) wtab.where[i] = 0;
  auto reduce = [&](int cls) -> int {
    wgrad_tc_reduce_kernel<<<grid_for(total, 256), 256, 0, st>>>(
        (const float*)workspace, q->dw, q->Ca, q->Cb, pl.u_is_x, a.s_chunks, s0.ksplit, s0.u_tiles * a.s_chunks * a.nroles,
        q->accumulate, cls, a.CU, a.pair_ok, seg1_tile, s1.ksplit, s1.u_tiles * a.s_chunks * a.nroles, (int64_t)s1.offset_floats,
        upconv ? 1 : 0, wtab, a.nroles, a.kw0, q->dw_ld, q->dw_c0, q->dw_transposed, g_reduce_s_fastest);
    MEDNET_LAUNCH_CHECK();
    return MEDNET_OK;
  };
  if (!convt) {
    int r = launch_all();
    if (r != MEDNET_OK) return r;
    r = reduce(-1);
    if (r != MEDNET_OK) return r;
  } else {
    // one pass per parity class of dY: window taps {1} (parity 0) / {0, 1} (parity 1) per axis -> at most 4 of the 9
    // (kd, kh) groups (mirrored when U is the strided operand); the reduce pass of each class writes only its taps
    for (int cls = 0; cls < 8; ++cls) {
      uint32_t gm = 0;
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh) {
          if (upconv) {    // dY parity 0: window taps {1, 2}, parity 1: {0, 1} (offset of the dY sub-grid = -c)
            if (((cls & 4) ? kd > 1 : kd < 1) || ((cls & 2) ? kh > 1 : kh < 1)) continue;
          } else {         // transposed conv: window tap 1 always, tap 0 for parity 1 only
            if (kd == 2 || kh == 2 || (kd == 0 && !(cls & 4)) || (kh == 0 && !(cls & 2))) continue;
          }
          const int gd = pl.u_is_x ? 2 - kd : kd, gh = pl.u_is_x ? 2 - kh : kh;
          gm |= 1u << (gd * 3 + gh);
        }
      a.cls = cls; a.gmask = gm;
      {                  // the (1, 2 or 4) needed groups, all in role 0 (A/B switch off: two per role)
        a.use_tab = 1; wtab.use = 1;
        const int per_role = a.nroles == 1 ? 4 : 2;
        int cnt = 0;
        for (int r = 0; r < 2; ++r)
          for (int i = 0; i < 5; ++i) a.tab[r][i] = -1;
        for (int g = 0; g < 9; ++g)
          if ((gm >> g) & 1u) {
            const int r = cnt / per_role, sl = cnt % per_role;
            a.tab[r][sl] = (signed char)g;
            wtab.where[g] = (signed char)(r * 8 + sl);
            ++cnt;
          }
      }
      if (g_class_merge) {          // kw windows of this class (same rule as kd, kh above), mirrored when U is the strided operand
        int lo = 3, hi = -1;
        for (int kw = 0; kw < 3; ++kw) {
          if (upconv ? ((cls & 1) ? kw > 1 : kw < 1) : (kw == 2 || (kw == 0 && !(cls & 1)))) continue;
          const int gw = pl.u_is_x ? 2 - kw : kw;
          if (gw < lo) lo = gw;
          if (gw > hi) hi = gw;
        }
        a.kw0 = lo;
        a.ncols = (hi - lo + 1) * CS;
      }
      int r = launch_all();
      if (r != MEDNET_OK) return r;
      r = reduce(cls);
      if (r != MEDNET_OK) return r;
    }
  }
  if (q->dbias != nullptr) {
    void* cpart = (char*)workspace + pl.partial_bytes;
    if (convt)    // bias gradient = column sums of the output-gradient operand (b for the transposed conv)
      return colsum_bias(q->b, q->dtype, (int64_t)q->N * q->Db * q->Hb * q->Wb, q->Cb, q->dbias, q->accumulate, cpart, st);
    return colsum_bias(q->a, q->dtype, (int64_t)q->N * q->Da * q->Ha * q->Wa, q->Ca, q->dbias, q->accumulate, cpart, st);
  }
  return MEDNET_OK;
}

}  // namespace mednet
