// b200sv tile engine: whole runs of 1- and 2-qubit gates per HBM pass (sm_100a).
//
// One HBM pass applies a *sequence* of 1-/2-qubit gates (and, for batched noisy shots, per-state Pauli ops) whose
// qubits fit a 12-bit tile: 2^12 amplitudes (64 KiB) are staged in shared memory with cp.async (LDGSTS.128, no
// register staging), the gates run as "rounds" -- each thread pulls the 16 amplitudes of a 4-bit sub-block into
// registers, applies the round's gates there (16 DFMA per amplitude per 2-qubit gate instead of 4*2^k for a fused
// dense block), writes them back -- and the tile is streamed out again.  HBM traffic per pass stays 2*16*2^n bytes
// while ~9 gates ride on it, which moves QV-style circuits from ~2 gates per pass (dense k<=4 fusion) past the
// FP64/HBM balance point (~90 DFMA per amplitude per pass on B200): the passes are FP64 bound.
//
// File map (top to bottom):
//   device   round code (apply2 / apply1 / diagonal / Pauli forms, run_rounds<MODE>), tile_pass_kernel (one tile per
//            CTA), tile_pipe_kernel (three tile buffers per SM), run_rounds_f32 (float2 slot pairs), tile_pipe2_kernel
//            (16 compute warps + 4 memory warps, mbarrier hand-offs: the default for dense-gate and noisy passes)
//   host     emulate_tile_pass (CPU interpreter of a parameter block: scheduler self-test), build_round /
//            plan_segments (swizzle-aware lane choice, warp-local segments), build_slot_rounds (fast rounds, folded
//            Paulis), run_tile_pass / run_tile_pass_f32, PassPacker (ready-set pass selection),
//            absorb_one_qubit_gates (queue-level fusion), apply_gate_sequence (entry point of the C ABI calls)
// DESIGN.md "The tile engine" has the measurements behind each of these pieces.
//
// Reference counterpart: the blocked-gate queue of the Thrust path
// (chunk/device_chunk_container.hpp:999-1108 queue_blocked_gate,
//  :1208 dev_apply_shared_memory_blocked_gates: <= 64 one-qubit gates on <= 10
//  "blocked" qubits per launch), which no in-tree pass ever enables
// (SURVEY Appendix B).  Here the queue is general (2-qubit gates, any qubits)
// and the scheduler below decides passes and rounds.
//
// Shared-memory layout: tile-local index j (bit u of j <-> global bit tb[u],
// tb sorted ascending and always containing the low global bits so that global
// accesses stay in >= 128 B runs) is stored at 16-byte slot  j ^ S(j)  with
//   S(j) = XOR_{u >= 3, bit u of j set} v[u],  v = {.,.,., 1,2,4, 3,6,5, 7, 1, 2}
// a GF(2)-linear swizzle chosen so that for ANY 4 round positions three of the
// remaining eight positions have linearly independent bank vectors: the host
// maps lane bits 0..2 to those, which makes every quarter-warp LDS.128/STS.128
// of a round hit 8 distinct 16-byte bank groups (conflict free), and keeps the
// staging accesses (consecutive j) conflict free as well.  Because the swizzle
// is linear, addresses are  phys(base) ^ phys(offset): one XOR per access.
#include "common.cuh"
#include <array>
#include <complex>

namespace b200sv {

constexpr int kMaxTB = 12;            // tile bits: 12 (64 KiB, 256 threads, 2 CTAs/SM) or 11 (32 KiB, 128 threads, 4 CTAs/SM)
constexpr int kHiCount = 16;          // staging iterations per thread = 2^TB / threads (one 16-amp group per thread)
constexpr int kRoundBits = 4;
constexpr int kMaxRounds = 64;     // gate rounds + one round per sampled-noise Pauli op
constexpr int kMaxTileGates = 16;   // 4x4 matrices per pass (kernel-parameter space)
constexpr int kMaxRoundGates = 12;  // dense gates + per-state Pauli ops in one round

__host__ __device__ constexpr int swz_vec(int u) {
  return u < 3 ? (1 << u) : u == 3 ? 1 : u == 4 ? 2 : u == 5 ? 4 : u == 6 ? 3 : u == 7 ? 6 : u == 8 ? 5 : u == 9 ? 7
                                                                                      : u == 10 ? 1 : 2;
}
__host__ __device__ inline uint32_t phys_slot(uint32_t j) {
  uint32_t s = 0;
  for (int u = 3; u < kMaxTB; u++)
    if ((j >> u) & 1u) s ^= (uint32_t)swz_vec(u);
  return j ^ s;
}

struct TileRound {
  uint16_t eoff[16];            // phys(sum_i bit_i(e) << pos[i]) for the 16 elements of a sub-block
  uint16_t eoff_ld[16];         // slot rounds of noisy passes load through this copy: eoff with the round's bare cx
                                // gates folded in as a permutation of the elements (== eoff without cx slots)
  uint16_t gbit[8];             // phys(1 << tpos[i]): contribution of group-id bit i
  uint8_t ngates;
  uint8_t npre;                 // number of valid entries in pre[]
  uint8_t pre[4];               // slot rounds: staged-code index of a per-state Pauli applied to round bit i BEFORE the
                                // round's gates (sampled noise folded into the next gate round on that qubit), or kMaxRounds = none
  uint8_t npre2;                // number of valid entries in pre2[]
  uint8_t pre2[4];              // like pre[], but applied AFTER the round's folded cx gates (through eoff_ld): the noise
                                // that sits between a folded cx and the 1-qubit gates of its slot
  uint8_t sync;                 // 1: CTA barrier after this round; 0: the next round stays inside each warp's sub-tile
  uint8_t fast;                 // 2 / 1: exactly two / one dense 4x4 block(s), on round bits (0,1) [and (2,3)]:
                                // straight-line code (LDS, DFMA and STS interleave, no form dispatch); 5: one per-state
                                // Pauli on round bit 0 (gate[0] = code slot); 0: generic
  uint8_t form[kMaxRoundGates];  // 0..5: 2-qubit on round-bit pair; 6..9: 1-qubit on round bit (form-6);
                                 // 10..13: per-state Pauli on round bit (form-10), code table slot in gate[];
                                 // 14..19: DIAGONAL 2-qubit on round-bit pair (form-14 as 0..5); 20..23: diagonal 1-qubit
  uint16_t gate[kMaxRoundGates]; // index into mats (dense) or error-code slot (Pauli)
};
struct TilePassParams {
  double2 mats[kMaxTileGates][16];  // 2q: row-major 4x4 with matrix bit0 <-> lower round bit; 1q: first 4 entries
  uint64_t goff_hi[kHiCount];       // global offset of tile-local bits kLoBits..11 (index m = j >> kLoBits)
  uint64_t goff_lo[8];              // global offset of tile-local bit u < kLoBits
  uint64_t ntiles;
  uint16_t soff_hi[kHiCount];       // phys(m << kLoBits)
  InsertList ins;                   // sorted tile bits (global positions)
  const uint8_t *codes;             // [slot][state] Pauli codes 0..3 = I,X,Y,Z (batched noisy shots), or null
  uint64_t nstates;
  int state_shift;                  // tile index >> state_shift = state (tile bits are all < num_qubits)
  int nrounds;
  int npauli;                       // Pauli rounds of a slot-round pass read their codes from a per-tile shared copy:
  uint16_t pauli_slot[kMaxRounds];  //   entry i = code-table slot of the pass's i-th Pauli op
  TileRound rounds[kMaxRounds];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// same, ordered against the surrounding shared-memory accesses of the issuing thread (buffer refill right after read-out)
__device__ __forceinline__ void cp_async16_ordered(void *smem, const void *gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// 2-qubit gate on round bits P0 < P1 of a 16-amplitude register block
template <int P0, int P1>
__device__ __forceinline__ void apply2(double2 (&a)[16], const double2 *__restrict__ m) {
  double2 mm[16];
#pragma unroll
  for (int i = 0; i < 16; i++) mm[i] = m[i];
#pragma unroll
  for (int o = 0; o < 4; o++) {
    // spread the two bits of o over the positions that are not P0 / P1
    int base = 0, ob = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
      if (b != P0 && b != P1) {
        if ((o >> ob) & 1) base |= 1 << b;
        ob++;
      }
    const int i0 = base, i1 = base | (1 << P0), i2 = base | (1 << P1), i3 = base | (1 << P0) | (1 << P1);
    const double2 x0 = a[i0], x1 = a[i1], x2 = a[i2], x3 = a[i3];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      double2 acc = mk<double>(0, 0);
      cfma(acc, mm[r * 4 + 0], x0);
      cfma(acc, mm[r * 4 + 1], x1);
      cfma(acc, mm[r * 4 + 2], x2);
      cfma(acc, mm[r * 4 + 3], x3);
      a[r == 0 ? i0 : r == 1 ? i1 : r == 2 ? i2 : i3] = acc;
    }
  }
}
template <int P>
__device__ __forceinline__ void apply1(double2 (&a)[16], const double2 *__restrict__ m) {
  const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    if (i & (1 << P)) continue;
    const double2 x0 = a[i], x1 = a[i | (1 << P)];
    double2 y0 = mk<double>(0, 0), y1 = mk<double>(0, 0);
    cfma(y0, m00, x0); cfma(y0, m01, x1);
    cfma(y1, m10, x0); cfma(y1, m11, x1);
    a[i] = y0;
    a[i | (1 << P)] = y1;
  }
}

// diagonal gates (cz / cp / rzz / phase ...): one complex multiply per amplitude instead of 4 (2-qubit) or 2 (1-qubit)
// complex FMAs.  m[0..3] (m[0..1]) = the diagonal, index = bit(P0) + 2 bit(P1).
template <int P0, int P1>
__device__ __forceinline__ void apply_diag2(double2 (&a)[16], const double2 *__restrict__ m) {
  const double2 d0 = m[0], d1 = m[1], d2 = m[2], d3 = m[3];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int idx = ((i >> P0) & 1) | (((i >> P1) & 1) << 1);
    const double2 d = idx == 0 ? d0 : idx == 1 ? d1 : idx == 2 ? d2 : d3;
    const double2 x = a[i];
    a[i] = mk<double>(d.x * x.x - d.y * x.y, d.x * x.y + d.y * x.x);
  }
}
template <int P>
__device__ __forceinline__ void apply_diag1(double2 (&a)[16], const double2 *__restrict__ m) {
  const double2 d0 = m[0], d1 = m[1];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const double2 d = ((i >> P) & 1) ? d1 : d0;
    const double2 x = a[i];
    a[i] = mk<double>(d.x * x.x - d.y * x.y, d.x * x.y + d.y * x.x);
  }
}

// per-state Pauli on round bit P (apply_pauli semantics, qubitvector.hpp:2393-2437: swap, Z sign, (-i)^num_y):
// pure moves and sign flips.  `code` is uniform over the CTA (a tile never straddles two states).
template <int P>
__device__ __forceinline__ void apply_pauli_reg(double2 (&a)[16], int code) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    if (i & (1 << P)) continue;
    const int j = i | (1 << P);
    const double2 x0 = a[i], x1 = a[j];
    if (code == 1) { a[i] = x1; a[j] = x0; }
    else if (code == 2) { a[i] = mk<double>(x1.y, -x1.x); a[j] = mk<double>(-x0.y, x0.x); }
    else { a[j] = mk<double>(-x1.x, -x1.y); }
  }
}

// Bare cx gates in slot rounds (batched noisy passes).  Noisy circuits keep their cx gates bare -- the sampled Paulis
// around them stop the host from absorbing neighbouring 1-qubit gates -- and a cx only permutes the 16 amplitudes of a
// register block: the host folds that permutation into the round's LOAD offsets (TileRound::eoff_ld), the kernel just
// skips the slot's arithmetic.  Half of a noisy pass's gate slots then cost no DFMA and no instruction at all.
// kSlotCxLo / kSlotCxHi = cx controlled by the slot's lower / upper round bit; other forms = dense 4x4.
constexpr int kSlotCxLo = 24, kSlotCxHi = 25;
template <int P0, int P1>
__device__ __forceinline__ void slot_gate(double2 (&a)[16], const int form, const double2 *__restrict__ m) {
  if (form < kSlotCxLo) apply2<P0, P1>(a, m);  // `form` comes from the parameter block: uniform branch
}

// Sampled noise folded into a gate round: the Pauli codes (staged per tile in shared memory) of the ops that precede
// the round's gates on its four bits.  Identity draws (99 %) cost four shared-memory bytes and a vote; a hit takes the
// block through registers once more (same thread, same slots: no synchronisation) before the straight-line gate code,
// which stays untouched.  pre index kMaxRounds ("none") reads a constant 0.
__device__ __forceinline__ void apply_pre_paulis(double2 *__restrict__ tile, const uint32_t base, const uint16_t (&eoff)[16],
                                                 const uint8_t (&pre)[4], const uint8_t *scodes, const bool valid) {
  const int c0 = __shfl_sync(0xffffffffu, (int)scodes[pre[0]], 0);
  const int c1 = __shfl_sync(0xffffffffu, (int)scodes[pre[1]], 0);
  const int c2 = __shfl_sync(0xffffffffu, (int)scodes[pre[2]], 0);
  const int c3 = __shfl_sync(0xffffffffu, (int)scodes[pre[3]], 0);
  if (c0 | c1 | c2 | c3) {
    // rare path (a few per cent of the rounds): pair by pair in shared memory, runtime bit and code, so that it adds
    // almost no registers or code next to the straight-line gate blocks
#pragma unroll 1
    for (int b = 0; b < 4; b++) {
      const int code = b == 0 ? c0 : b == 1 ? c1 : b == 2 ? c2 : c3;
      if (!code) continue;
#pragma unroll 1
      for (int j = 0; j < 8; j++) {
        const int i0 = ((j >> b) << (b + 1)) | (j & ((1 << b) - 1)), i1 = i0 | (1 << b);
        double2 *s0 = &tile[base ^ eoff[i0]], *s1 = &tile[base ^ eoff[i1]];
        const double2 x0 = *s0, x1 = *s1;
        if (!valid) continue;
        if (code == 1) { *s0 = x1; *s1 = x0; }
        else if (code == 2) { *s0 = mk<double>(x1.y, -x1.x); *s1 = mk<double>(-x0.y, x0.x); }
        else *s1 = mk<double>(-x1.x, -x1.y);
      }
    }
  }
}

// Pauli codes of one tile's state -> shared memory, by ONE warp with one (two) parallel loads: lane i fetches the code
// of the pass's i-th (and i+32-th) Pauli op.  The slot numbers come out of the parameter block with a UNIFORM index (a
// thread-indexed parameter read makes the compiler copy the block to local memory and takes the gate matrices off the
// uniform datapath), each lane keeping its own.  (The first version fetched the codes one after the other: ~40 dependent
// L2 round trips per tile, which the rest of the group sat out at the barrier -- half of a noisy pass's time.)
__device__ __forceinline__ void stage_pauli_codes(uint8_t *sc, const TilePassParams &p, const uint64_t state, const int lane) {
  uint32_t s0 = 0, s1 = 0;
  for (int i = 0; i < p.npauli; i++) {
    const uint32_t sl = p.pauli_slot[i];
    if ((i & 31) == lane) { if (i < 32) s0 = sl; else s1 = sl; }
  }
  if (lane < p.npauli) sc[lane] = p.codes[(size_t)s0 * p.nstates + state];
  if (lane + 32 < p.npauli) sc[lane + 32] = p.codes[(size_t)s1 * p.nstates + state];
  if (lane == 0) sc[kMaxRounds] = 0;  // "no Pauli"
}

// all rounds of one tile, in place in shared memory.  GROUPED = 0: the CTA is one 2^(TB-4)-thread group
// (__syncthreads); GROUPED = 1: 256-thread groups of a bigger CTA, named barrier 1 + grp.
// MODE 0: fast and generic rounds; 1: every round of the pass is fast; 2: generic code only (fast rounds carry
// generic forms too).  The pipelined kernel is instantiated per mode: with both code paths in one kernel ptxas
// runs out of uniform registers and demotes the generic path's matrix operands to vector registers.
// LASTSYNC = false: no barrier after the final round (the caller hands the tile over through an mbarrier that every
// thread arrives on, so the warps of a group may drift apart across tiles).
template <int GROUPED, int kLoBits, int MODE, bool LASTSYNC = true, int GT = 256>
__device__ __forceinline__ void run_rounds(double2 *__restrict__ tile, const int tid, const uint64_t t,
                                           const TilePassParams &p, const int grp, const bool valid,
                                           const uint8_t *scodes = nullptr) {
  for (int r = 0; r < p.nrounds; r++) {
    const TileRound &R = p.rounds[r];
    {
      const int g = tid;
      uint32_t base = 0;
#pragma unroll
      for (int i = 0; i < kLoBits; i++)
        if ((g >> i) & 1) base ^= R.gbit[i];
      if (MODE == 4 && R.npre) apply_pre_paulis(tile, base, R.eoff, R.pre, scodes, valid);
      if (MODE == 4 && R.npre2) apply_pre_paulis(tile, base, R.eoff_ld, R.pre2, scodes, valid);  // after the folded cx gates
      const int fast = MODE == 2 ? 0 : R.fast;  // MODE 4 = MODE 1 + Pauli rounds
      if (fast == 2) {
        double2 a[16];
#pragma unroll
        for (int e = 0; e < 16; e++) a[e] = tile[base ^ (MODE == 4 ? R.eoff_ld[e] : R.eoff[e])];
        if (MODE == 4) {
          slot_gate<0, 1>(a, R.form[0], p.mats[R.gate[0]]);
          slot_gate<2, 3>(a, R.form[1], p.mats[R.gate[1]]);
        } else {
          apply2<0, 1>(a, p.mats[R.gate[0]]);
          apply2<2, 3>(a, p.mats[R.gate[1]]);
        }
#pragma unroll
        for (int e = 0; e < 16; e++)
          if (valid) tile[base ^ R.eoff[e]] = a[e];
      } else if ((MODE == 0 || MODE == 4) && fast == 5) {
        // one per-state Pauli (sampled noise between gate rounds) on round bit 0, R.gate[0] = code-table slot.  Most
        // draws are the identity: the round then costs one byte load and a broadcast, no shared-memory traffic.
        int raw;
        if (MODE == 4) {
          const uint32_t sa = (uint32_t)__cvta_generic_to_shared(scodes) + R.gate[0];
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(raw) : "r"(sa));
        } else {
          raw = (int)p.codes[(size_t)p.pauli_slot[R.gate[0]] * p.nstates + (t >> p.state_shift)];
        }
        const int code = __shfl_sync(0xffffffffu, raw, 0);
        if (code) {
          double2 a[16];
#pragma unroll
          for (int e = 0; e < 16; e++) a[e] = tile[base ^ R.eoff[e]];
          apply_pauli_reg<0>(a, code);
#pragma unroll
          for (int e = 0; e < 16; e++)
            if (valid) tile[base ^ R.eoff[e]] = a[e];
        }
      } else if (MODE == 1 || MODE == 4 || fast == 1) {
        double2 a[16];
#pragma unroll
        for (int e = 0; e < 16; e++) a[e] = tile[base ^ (MODE == 4 ? R.eoff_ld[e] : R.eoff[e])];
        if (MODE == 4) slot_gate<0, 1>(a, R.form[0], p.mats[R.gate[0]]);
        else apply2<0, 1>(a, p.mats[R.gate[0]]);
#pragma unroll
        for (int e = 0; e < 16; e++)
          if (valid) tile[base ^ R.eoff[e]] = a[e];
      } else {
        double2 a[16];
#pragma unroll
        for (int e = 0; e < 16; e++) a[e] = tile[base ^ R.eoff[e]];
        for (int k = 0; k < R.ngates; k++) {
          const int form = R.form[k];
          if (form >= 10 && form < 14) {  // sampled noise: Pauli chosen per state (shot)
            // uniform over the CTA; the lane-0 broadcast lets the compiler see that (keeps the dense gates below in
            // convergent control flow, i.e. their matrices on the uniform datapath)
            const int code = __shfl_sync(0xffffffffu, (int)p.codes[(size_t)R.gate[k] * p.nstates + (t >> p.state_shift)], 0);
            if (code) {
              switch (form) {
              case 10: apply_pauli_reg<0>(a, code); break;
              case 11: apply_pauli_reg<1>(a, code); break;
              case 12: apply_pauli_reg<2>(a, code); break;
              default: apply_pauli_reg<3>(a, code); break;
              }
            }
            continue;
          }
          const double2 *m = p.mats[R.gate[k]];
          switch (form) {
          case 0: apply2<0, 1>(a, m); break;
          case 1: apply2<0, 2>(a, m); break;
          case 2: apply2<0, 3>(a, m); break;
          case 3: apply2<1, 2>(a, m); break;
          case 4: apply2<1, 3>(a, m); break;
          case 5: apply2<2, 3>(a, m); break;
          case 6: apply1<0>(a, m); break;
          case 7: apply1<1>(a, m); break;
          case 8: apply1<2>(a, m); break;
          case 9: apply1<3>(a, m); break;
          case 14: apply_diag2<0, 1>(a, m); break;
          case 15: apply_diag2<0, 2>(a, m); break;
          case 16: apply_diag2<0, 3>(a, m); break;
          case 17: apply_diag2<1, 2>(a, m); break;
          case 18: apply_diag2<1, 3>(a, m); break;
          case 19: apply_diag2<2, 3>(a, m); break;
          case 20: apply_diag1<0>(a, m); break;
          case 21: apply_diag1<1>(a, m); break;
          case 22: apply_diag1<2>(a, m); break;
          default: apply_diag1<3>(a, m); break;
          }
        }
#pragma unroll
        for (int e = 0; e < 16; e++)
          if (valid) tile[base ^ R.eoff[e]] = a[e];
      }
    }
    if (R.sync && (LASTSYNC || r + 1 < p.nrounds)) {
      if (GROUPED) {
        if (GT == 256) {
          if (grp) asm volatile("bar.sync 2, 256;" ::: "memory");
          else asm volatile("bar.sync 1, 256;" ::: "memory");
        } else {
          switch (grp) {  // constant barrier ids (a register id makes ptxas reserve all 16 barriers)
          case 0: asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory"); break;
          case 1: asm volatile("bar.sync 2, %0;" ::"n"(GT) : "memory"); break;
          case 2: asm volatile("bar.sync 3, %0;" ::"n"(GT) : "memory"); break;
          default: asm volatile("bar.sync 4, %0;" ::"n"(GT) : "memory"); break;
          }
        }
      } else __syncthreads();
    }
    else __syncwarp();
  }
}

#ifdef B200SV_TILE_PROFILE
__device__ unsigned long long g_tile_prof[4];  // load-wait, rounds, store, tiles (thread 0 of every CTA)
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(i, v) if (threadIdx.x == 0) atomicAdd(&g_tile_prof[i], (unsigned long long)(v))
#else
#define PROF_T(var)
#define PROF_ADD(i, v)
#endif

template <int TB>
__global__ void __launch_bounds__(1 << (TB - 4), TB == 12 ? 2 : 4)
tile_pass_kernel(double2 *__restrict__ psi, const __grid_constant__ TilePassParams p) {
  constexpr int kLoBits = TB - 4;  // tile-local bits covered by the thread id (= group id bits)
  extern __shared__ __align__(16) double2 tile[];
  const int tid = threadIdx.x;
  uint64_t glo = 0;
#pragma unroll
  for (int u = 0; u < kLoBits; u++)
    if ((tid >> u) & 1) glo |= p.goff_lo[u];
  const uint32_t slo = phys_slot((uint32_t)tid);

  for (uint64_t t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
    double2 *gt = psi + (insert_zeros(t, p.ins) | glo);
    PROF_T(c0);
#pragma unroll
    for (int m = 0; m < kHiCount; m++) cp_async16(&tile[slo ^ p.soff_hi[m]], gt + p.goff_hi[m]);
    cp_async_wait_all();
    __syncthreads();
    PROF_T(c1);

    run_rounds<0, kLoBits, 0>(tile, tid, t, p, 0, true);

    PROF_T(c2);
#pragma unroll
    for (int m = 0; m < kHiCount; m++) gt[p.goff_hi[m]] = tile[slo ^ p.soff_hi[m]];
    __syncthreads();
    PROF_T(c3);
    PROF_ADD(0, c1 - c0); PROF_ADD(1, c2 - c1); PROF_ADD(2, c3 - c2); PROF_ADD(3, 1);
  }
}

// ------------------------------------------------------------------------------------------ pipelined variant
// One 512-thread CTA per SM = two 256-thread groups that each run the load-wait / rounds / store sequence of
// tile_pass_kernel<12> on their own tile, plus a THIRD 64 KiB tile buffer: the buffer a group has just streamed
// out is refilled at once (same thread, same slots: LDS -> STG -> LDGSTS) with the tile the OTHER group will need
// two tiles later, so a tile's HBM latency runs behind the other group's rounds instead of stalling its consumer.
// Tile k of the CTA lives in buffer k % 3 and belongs to group k % 2; "full" mbarriers (256 cp.async arrivals)
// hand a loaded buffer across groups.  (Queueing view: 2 compute customers + 3 memory customers per SM instead of
// 2 + 2 -- profiles/r01_tile_phase_breakdown.md.)
constexpr int kPipeBufs = 3;
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n.reg .pred P1;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(a), "r"(parity) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1)
tile_pipe_kernel(double2 *__restrict__ psi, const __grid_constant__ TilePassParams p) {
  constexpr int kLoBits = 8;
  extern __shared__ __align__(16) double2 tiles[];  // kPipeBufs tiles, then the barriers and progress counters
  uint64_t *full = reinterpret_cast<uint64_t *>(tiles + kPipeBufs * 4096);
  volatile int *progress = reinterpret_cast<volatile int *>(full + kPipeBufs);  // [grp]: tiles whose rounds are done
  uint8_t *scodes = reinterpret_cast<uint8_t *>(full + kPipeBufs + 1);           // [grp][kMaxRounds + 16] Pauli codes of the tile's state
  // broadcast from lane 0: tells the compiler the group id is warp-uniform (keeps the gate matrices on the uniform datapath)
  const int grp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 8), 0), tid = threadIdx.x & 255;
  if (threadIdx.x == 0) {
    for (int b = 0; b < kPipeBufs; b++) mbar_init(&full[b], 256);
    progress[0] = 0;
    progress[1] = 0;
  }
  __syncthreads();
  uint64_t glo = 0;
#pragma unroll
  for (int u = 0; u < kLoBits; u++)
    if ((tid >> u) & 1) glo |= p.goff_lo[u];
  const uint32_t slo = phys_slot((uint32_t)tid);
  auto issue_load = [&](uint64_t k) {  // all 256 threads of one group; arrives on full[k % 3] when the copies land
    const uint64_t t = blockIdx.x + k * gridDim.x;
    double2 *buf = tiles + (k % kPipeBufs) * 4096;
    if (t < p.ntiles) {
      const double2 *gt = psi + (insert_zeros(t, p.ins) | glo);
#pragma unroll
      for (int m = 0; m < kHiCount; m++) cp_async16_ordered(&buf[slo ^ p.soff_hi[m]], gt + p.goff_hi[m]);
    }
    mbar_arrive_cp_async(&full[k % kPipeBufs]);
  };
  if (grp == 0) { issue_load(0); issue_load(2); }
  else issue_load(1);
  // Both groups run the same (uniform) number of iterations so that the rounds stay in convergent control flow
  // (gate matrices on the uniform datapath); a group without a tile in its last iteration computes on stale shared
  // memory with its stores predicated off.
  const int K = (int)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // tiles of this CTA
  for (int kk = 0; 2 * kk < K; kk++) {
    const int k = 2 * kk + grp;
    const bool valid = k < K;
    const uint64_t t = blockIdx.x + (uint64_t)k * gridDim.x;
    double2 *tile = tiles + (k % kPipeBufs) * 4096;
    if (valid) {
      // tile k >= 3 was requested by the other group after ITS tile k - 3: make sure that group is past its own wait
      // on this buffer before testing the barrier's phase parity (a waiter two phases ahead would read a stale parity)
      if (k >= 3) {
        const int need = ((k - 3) >> 1) + 1;
        while (progress[grp ^ 1] < need) { }
      }
      mbar_wait(&full[k % kPipeBufs], (uint32_t)((k / kPipeBufs) & 1));
    }
    if (MODE == 4) {  // this tile's state: one code per Pauli op of the pass, fetched once (not once per round)
      // the group's first warp fetches them (uniform branch, uniform parameter index: anything thread-indexed here makes
      // the compiler move the parameter block to local memory and the gate matrices off the uniform datapath)
      if (__shfl_sync(0xffffffffu, tid >> 5, 0) == 0)
        stage_pauli_codes(scodes + grp * (kMaxRounds + 16), p, valid ? (t >> p.state_shift) : 0, tid & 31);
      if (grp) asm volatile("bar.sync 2, 256;" ::: "memory");
      else asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    run_rounds<1, kLoBits, MODE>(tile, tid, t, p, grp, valid, scodes + grp * (kMaxRounds + 16));
    if (valid) {
      if (tid == 0) progress[grp] = kk + 1;
      double2 *gt = psi + (insert_zeros(t, p.ins) | glo);
#pragma unroll
      for (int m = 0; m < kHiCount; m++) gt[p.goff_hi[m]] = tile[slo ^ p.soff_hi[m]];
      issue_load(k + 3);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ single precision
// float2 amplitudes are 8 bytes: a 16-byte swizzle slot holds the two amplitudes that differ in global qubit 0, so
// the whole staging / buffer-rotation machinery above is reused unchanged on the state seen as 2^(n-1) 16-byte
// elements (tile = 2^12 slots = 2^13 amplitudes).  A thread's 16 slots are a 5-bit block: bit 0 = global qubit 0,
// bits 1..4 = the round's slot positions.  Rounds are straight-line only: gate A on bits (1,2) -- or (0,1) when it
// involves qubit 0 -- and optionally gate B on bits (3,4); the host promotes 1-qubit gates to 4x4.
// Blackwell packed FP32 (`fma.rn.f32x2`, SASS FFMA2): one instruction does the two FMAs of a complex-times-real step,
//   acc(re, im) += (m.re, m.re) * (x.re, x.im);   acc(re, im) += (-m.im, m.im) * (x.im, x.re)
// -- the same four roundings in the same order as cfma().  The host stores every matrix entry of a float pass as the two
// pairs {m.re, m.re, -m.im, m.im} (16 bytes, the room a double-precision entry takes), so that both multipliers are
// 64-bit uniform-register operands straight from LDCU; the swap of x is an operand modifier (R.F32x2.LO_HI).  A 4x4
// complex mat-vec is 32 FFMA2 instead of 64 FFMA.  The float rounds were issue bound (ncu: 87 % of the issue slots
// busy, 75 % of them FFMA): halving the FMA instructions is what moves them.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void cfma_f32x2(unsigned long long &acc, const ulonglong2 m, const float2 x) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(m.x), "l"(pack_f32x2(x.x, x.y)));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(m.y), "l"(pack_f32x2(x.y, x.x)));
}
template <int P0, int P1>
__device__ __forceinline__ void apply2f(float2 (&a)[32], const ulonglong2 *__restrict__ m) {
  ulonglong2 mm[16];
#pragma unroll
  for (int i = 0; i < 16; i++) mm[i] = m[i];
#pragma unroll
  for (int o = 0; o < 8; o++) {
    int base = 0, ob = 0;
#pragma unroll
    for (int b = 0; b < 5; b++)
      if (b != P0 && b != P1) {
        if ((o >> ob) & 1) base |= 1 << b;
        ob++;
      }
    const int i0 = base, i1 = base | (1 << P0), i2 = base | (1 << P1), i3 = base | (1 << P0) | (1 << P1);
    const float2 x0 = a[i0], x1 = a[i1], x2 = a[i2], x3 = a[i3];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      unsigned long long acc2 = 0ull;
      cfma_f32x2(acc2, mm[r * 4 + 0], x0);
      cfma_f32x2(acc2, mm[r * 4 + 1], x1);
      cfma_f32x2(acc2, mm[r * 4 + 2], x2);
      cfma_f32x2(acc2, mm[r * 4 + 3], x3);
      float2 acc;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(acc2));
      a[r == 0 ? i0 : r == 1 ? i1 : r == 2 ? i2 : i3] = acc;
    }
  }
}

// R.fast: 1 = A on bits (1,2); 2 = A on (1,2) + B on (3,4); 3 = A on (0,1); 4 = A on (0,1) + B on (3,4)
__device__ __forceinline__ void run_rounds_f32(double2 *__restrict__ tile, const int tid, const TilePassParams &p,
                                               const int grp, const bool valid) {
  for (int r = 0; r < p.nrounds; r++) {
    const TileRound &R = p.rounds[r];
    uint32_t base = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
      if ((tid >> i) & 1) base ^= R.gbit[i];
    const ulonglong2 *mA = reinterpret_cast<const ulonglong2 *>(p.mats[R.gate[0]]);  // {re, re, -im, im} per entry
    const ulonglong2 *mB = reinterpret_cast<const ulonglong2 *>(p.mats[R.gate[1]]);
    const int kind = R.fast;
    float2 a[32];
#define B200_F32_LOAD                                                                  \
  _Pragma("unroll") for (int e = 0; e < 16; e++) {                                     \
    const float4 v = *reinterpret_cast<const float4 *>(&tile[base ^ R.eoff[e]]);       \
    a[2 * e] = make_float2(v.x, v.y);                                                  \
    a[2 * e + 1] = make_float2(v.z, v.w);                                              \
  }
#define B200_F32_STORE                                                                                           \
  _Pragma("unroll") for (int e = 0; e < 16; e++) if (valid)                                                      \
      *reinterpret_cast<float4 *>(&tile[base ^ R.eoff[e]]) = make_float4(a[2 * e].x, a[2 * e].y, a[2 * e + 1].x, a[2 * e + 1].y);
    if (kind == 2) {
      B200_F32_LOAD
      apply2f<1, 2>(a, mA);
      apply2f<3, 4>(a, mB);
      B200_F32_STORE
    } else if (kind == 1) {
      B200_F32_LOAD
      apply2f<1, 2>(a, mA);
      B200_F32_STORE
    } else if (kind == 4) {
      B200_F32_LOAD
      apply2f<0, 1>(a, mA);
      apply2f<3, 4>(a, mB);
      B200_F32_STORE
    } else {
      B200_F32_LOAD
      apply2f<0, 1>(a, mA);
      B200_F32_STORE
    }
#undef B200_F32_LOAD
#undef B200_F32_STORE
    if (R.sync && r + 1 < p.nrounds) {
      if (grp) asm volatile("bar.sync 2, 256;" ::: "memory");
      else asm volatile("bar.sync 1, 256;" ::: "memory");
    } else __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ pipelined variant 2
// Same three-buffer rotation, but the tile traffic is taken off the compute groups altogether: four extra warps
// (threads 512..639; registers are handed out to warps four at a time, so 20 warps x 96 registers is the fit) stream every finished tile out and the tile three positions later in, so the 16 compute warps
// only ever wait on barriers, read/write shared memory and issue DFMA.  Hand-offs: full[b] (128 cp.async arrivals)
// memory -> compute, done[b] (256 arrivals) compute -> memory.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// MODE 1: double precision, fast rounds; MODE 3: single precision (psi = the state as 16-byte slots)
// TB = 11: 2^11-amplitude tiles (32 KiB), FOUR compute groups of 128 threads and six buffers: every scheduler of the SM
// then holds one warp of each group, four independent phases (shared-memory reads / DFMA block / writes / barrier)
// instead of two pairs in lock step.  The smaller tile carries fewer gates per pass, which is affordable: the passes
// are FP64 bound, their time follows the gate count, not the pass count (B200SV_TILE_PIPE=3).
template <int MODE, int TB = 12>
__global__ void __maxnreg__(96) tile_pipe2_kernel(double2 *__restrict__ psi, const __grid_constant__ TilePassParams p) {
  constexpr int kLoBits = TB - 4, kTile = 1 << TB, kGT = 1 << (TB - 4), kGroups = 512 / kGT, kBufs = TB == 12 ? kPipeBufs : 6;
  constexpr int kMemIters = kTile / 128;
  extern __shared__ __align__(16) double2 tiles[];
  uint64_t *full = reinterpret_cast<uint64_t *>(tiles + kBufs * kTile);
  uint64_t *done = full + kBufs;
  uint8_t *scodes = reinterpret_cast<uint8_t *>(done + kBufs);  // MODE 4: [grp][2][kMaxRounds + 16] Pauli codes
  // lane-0 broadcasts: warp-uniform role / group ids the compiler can see (uniform branches, uniform datapath)
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0)
    for (int b = 0; b < kBufs; b++) { mbar_init(&full[b], 128); mbar_init(&done[b], kGT); }
  __syncthreads();
  const int K = (int)((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // tiles of this CTA
  if (wid >= 16) {
    // ---- memory warps: thread mt moves tile-local amplitudes j = mt + 128 i
    const int mt = threadIdx.x - 512;
    uint64_t glo = 0;
#pragma unroll
    for (int u = 0; u < 7; u++)
      if ((mt >> u) & 1) glo |= p.goff_lo[u];
    const uint32_t slo = phys_slot((uint32_t)mt);
    // tile-local bit 7 belongs to the staging-low bits for 2^12 tiles (kLoBits = 8) and to the high index for 2^11 ones
    auto goff = [&](int i) { return TB == 12 ? (((i & 1) ? p.goff_lo[7] : 0) | p.goff_hi[i >> 1]) : p.goff_hi[i]; };
    auto soff = [&](int i) { return TB == 12 ? (phys_slot((uint32_t)(i & 1) << 7) ^ (uint32_t)p.soff_hi[i >> 1]) : (uint32_t)p.soff_hi[i]; };
    auto load = [&](int k) {
      double2 *buf = tiles + (k % kBufs) * kTile;
      const double2 *gt = psi + (insert_zeros(blockIdx.x + (uint64_t)k * gridDim.x, p.ins) | glo);
#pragma unroll 16
      for (int i = 0; i < kMemIters; i++) cp_async16_ordered(&buf[slo ^ soff(i)], gt + goff(i));
      mbar_arrive_cp_async(&full[k % kBufs]);
    };
    for (int k = 0; k < kBufs && k < K; k++) load(k);
    for (int k = 0; k < K; k++) {
      mbar_wait(&done[k % kBufs], (uint32_t)((k / kBufs) & 1));
      double2 *buf = tiles + (k % kBufs) * kTile;
      double2 *gt = psi + (insert_zeros(blockIdx.x + (uint64_t)k * gridDim.x, p.ins) | glo);
#pragma unroll 16
      for (int i = 0; i < kMemIters; i++) gt[goff(i)] = buf[slo ^ soff(i)];
      if (k + kBufs < K) load(k + kBufs);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }
  // ---- compute groups
  const int grp = wid >> (TB - 9), tid = threadIdx.x & (kGT - 1);
  for (int kk = 0; kGroups * kk < K; kk++) {
    const int k = kGroups * kk + grp;
    const bool valid = k < K;
    const uint64_t t = blockIdx.x + (uint64_t)k * gridDim.x;
    double2 *tile = tiles + (k % kBufs) * kTile;
    if (valid) {
      // order the parity tests: tile k - 3 (other group) left this buffer => full[] is in tile k's phase or past it
      if (k >= kBufs) mbar_wait(&done[k % kBufs], (uint32_t)(((k - kBufs) / kBufs) & 1));
      mbar_wait(&full[k % kBufs], (uint32_t)((k / kBufs) & 1));
    }
    // MODE 4: Pauli codes of this tile's state, staged by the group's first warp (see tile_pipe_kernel).  Two copies per
    // group, alternating by tile: there is no barrier at the end of a tile here, so the first warp may already be staging
    // the group's next tile while the others still read this one's codes (the barrier below bounds the lead to one tile).
    uint8_t *sc = scodes + (grp * 2 + (kk & 1)) * (kMaxRounds + 16);
    if (MODE == 4) {
      if (__shfl_sync(0xffffffffu, tid >> 5, 0) == 0) stage_pauli_codes(sc, p, valid ? (t >> p.state_shift) : 0, tid & 31);
      if (grp) asm volatile("bar.sync 2, 256;" ::: "memory");
      else asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (MODE == 3) run_rounds_f32(tile, tid, p, grp, valid);
    else run_rounds<1, kLoBits, MODE == 3 ? 1 : MODE, false, kGT>(tile, tid, t, p, grp, valid, sc);
    if (valid) mbar_arrive(&done[k % kBufs]);  // release: this thread's shared-memory writes are visible to the waiter
  }
}

// ------------------------------------------------------------------------------------------ host scheduler
typedef std::complex<double> cd_t;
struct QGate {
  int nq;
  int q[2];
  const double *mat;  // column-major complex<double>, 4 or 16 entries; null for a per-state Pauli
  int slot;           // error-code table slot for a per-state Pauli (kind 3)
};

static bool is_diag(const QGate &g) {
  if (!g.mat) return false;
  const int dim = 1 << g.nq;
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++)
      if (i != j && (g.mat[2 * (i + dim * j)] != 0.0 || g.mat[2 * (i + dim * j) + 1] != 0.0)) return false;
  return true;
}
static uint64_t qmask(const QGate &g) { return (1ull << g.q[0]) | (g.nq == 2 ? (1ull << g.q[1]) : 0); }

// choose 3 group-id "lane" positions with independent swizzle vectors among the non-round positions
static bool pick_lane_positions(const std::vector<int> &free_pos, int out[3]) {
  const int n = (int)free_pos.size();
  for (int a = 0; a < n; a++)
    for (int b = a + 1; b < n; b++)
      for (int c = b + 1; c < n; c++) {
        const int va = swz_vec(free_pos[a]), vb = swz_vec(free_pos[b]), vc = swz_vec(free_pos[c]);
        if (va != vb && va != vc && vb != vc && (va ^ vb) != vc) {
          out[0] = free_pos[a]; out[1] = free_pos[b]; out[2] = free_pos[c];
          return true;
        }
      }
  return false;
}

// ------------------------------------------------------------------------------------------ scheduler self-test
// Host interpreter of a TilePassParams block: follows the kernels' addressing step by step (staging through
// insert_zeros / goff_* into swizzled slots, per-thread round blocks through gbit / eoff, fast kinds and generic
// forms, store back) on a HOST array.  It exists so that the pass / round / segment scheduler -- host code that
// otherwise only runs in front of a GPU -- is covered by the CPU test-suite (tests/test_tile_scheduler.py); it also
// checks the invariant the warp-local segments rely on: between two rounds that are separated by __syncwarp() only,
// no slot changes hands between warps.  Never reachable from a handle.
template <typename T>
static void emulate_tile_pass(const TilePassParams &p, void *host, const uint8_t *codes, bool f32_layout) {
  typedef std::complex<T> C;
  const int kSlots = 4096, kPer = f32_layout ? 2 : 1;  // amplitudes per 16-byte slot
  C *psi = reinterpret_cast<C *>(host);
  std::vector<C> tile((size_t)kSlots * kPer);
  std::vector<int> owner(kSlots);
  for (uint64_t t = 0; t < p.ntiles; t++) {
    const uint64_t tbase = insert_zeros(t, p.ins);
    // stage in with the memory-warp mapping (thread mt moves tile-local slots mt + 128 i) ...
    for (int mt = 0; mt < 128; mt++) {
      uint64_t glo = 0;
      for (int u = 0; u < 7; u++)
        if ((mt >> u) & 1) glo |= p.goff_lo[u];
      const uint32_t slo = phys_slot((uint32_t)mt);
      for (int i = 0; i < 32; i++) {
        const uint64_t g = (tbase | glo) + (((i & 1) ? p.goff_lo[7] : 0) | p.goff_hi[i >> 1]);
        const uint32_t sl = slo ^ phys_slot((uint32_t)(i & 1) << 7) ^ (uint32_t)p.soff_hi[i >> 1];
        for (int c = 0; c < kPer; c++) tile[(size_t)sl * kPer + c] = psi[g * kPer + c];
      }
    }
    std::fill(owner.begin(), owner.end(), -1);
    for (int r = 0; r < p.nrounds; r++) {
      const TileRound &R = p.rounds[r];
      for (int tid = 0; tid < 256; tid++) {
        uint32_t base = 0;
        for (int i = 0; i < 8; i++)
          if ((tid >> i) & 1) base ^= R.gbit[i];
        const int nb = f32_layout ? 5 : 4, na = 1 << nb;
        C a[32];
        for (int e = 0; e < 16; e++) {
          const uint32_t sl = base ^ R.eoff[e];
          if (sl >= (uint32_t)kSlots) throw Error("selftest: slot out of range");
          if (owner[sl] >= 0 && owner[sl] != (tid >> 5)) throw Error("selftest: a slot changes warps inside a warp-local segment");
          owner[sl] = tid >> 5;
          for (int c = 0; c < kPer; c++) a[e * kPer + c] = tile[(size_t)sl * kPer + c];
        }
        auto dense2 = [&](int P0, int P1, const C *M) {  // M row-major 4x4, index bit 0 <-> P0
          for (int i = 0; i < na; i++) {
            if (i & ((1 << P0) | (1 << P1))) continue;
            const int idx[4] = {i, i | (1 << P0), i | (1 << P1), i | (1 << P0) | (1 << P1)};
            C x[4], y[4];
            for (int c = 0; c < 4; c++) x[c] = a[idx[c]];
            for (int rr = 0; rr < 4; rr++) {
              y[rr] = 0;
              for (int c = 0; c < 4; c++) y[rr] += M[rr * 4 + c] * x[c];
            }
            for (int c = 0; c < 4; c++) a[idx[c]] = y[c];
          }
        };
        if (f32_layout) {
          std::complex<float> MA[16], MB[16];
          const float4 *mA = reinterpret_cast<const float4 *>(p.mats[R.gate[0]]);  // {re, re, -im, im} per entry
          const float4 *mB = reinterpret_cast<const float4 *>(p.mats[R.gate[1]]);
          for (int i = 0; i < 16; i++) {
            if (mA[i].x != mA[i].y || mA[i].z != -mA[i].w || mB[i].x != mB[i].y || mB[i].z != -mB[i].w)
              throw Error("selftest: float matrix entry is not stored as {re, re, -im, im}");
            MA[i] = {mA[i].x, mA[i].w};
            MB[i] = {mB[i].x, mB[i].w};
          }
          const C *A = reinterpret_cast<const C *>(MA), *B = reinterpret_cast<const C *>(MB);
          if (R.fast == 1 || R.fast == 2) dense2(1, 2, A); else dense2(0, 1, A);
          if (R.fast == 2 || R.fast == 4) dense2(3, 4, B);
        } else {
          static const int pair_of[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
          if (R.fast == 1 || R.fast == 2) {
            for (int P = 0; P < 4; P++) {
              if (R.pre[P] == kMaxRounds) continue;
              const int code = codes[(size_t)p.pauli_slot[R.pre[P]] * p.nstates + (t >> p.state_shift)];
              for (int i = 0; i < na && code; i++) {
                if (i & (1 << P)) continue;
                const int j = i | (1 << P);
                const C x0 = a[i], x1 = a[j];
                if (code == 1) { a[i] = x1; a[j] = x0; }
                else if (code == 2) { a[i] = C(x1.imag(), -x1.real()); a[j] = C(-x0.imag(), x0.real()); }
                else a[j] = -x1;
              }
            }
          }
          if (R.fast == 1 || R.fast == 2) {
            // the kernels of noisy passes load the block through eoff_ld (bare cx slots folded in as a permutation of the
            // elements) AFTER the pre-Paulis went through shared memory: replay exactly that table
            C b[16];
            for (int e = 0; e < 16; e++) {
              int src = -1;
              for (int x = 0; x < 16; x++)
                if (R.eoff[x] == R.eoff_ld[e]) src = x;
              if (src < 0) throw Error("selftest: eoff_ld is not a permutation of eoff");
              b[e] = a[src];
            }
            std::copy(b, b + 16, a);
            for (int P = 0; P < 4; P++) {  // noise between a folded cx and the gates of its slot
              if (R.pre2[P] == kMaxRounds) continue;
              const int code = codes[(size_t)p.pauli_slot[R.pre2[P]] * p.nstates + (t >> p.state_shift)];
              for (int i = 0; i < na && code; i++) {
                if (i & (1 << P)) continue;
                const int j = i | (1 << P);
                const C x0 = a[i], x1 = a[j];
                if (code == 1) { a[i] = x1; a[j] = x0; }
                else if (code == 2) { a[i] = C(x1.imag(), -x1.real()); a[j] = C(-x0.imag(), x0.real()); }
                else a[j] = -x1;
              }
            }
          }
          if (R.fast == 5) {
            const int code = codes[(size_t)p.pauli_slot[R.gate[0]] * p.nstates + (t >> p.state_shift)];
            for (int i = 0; i < na && code; i++) {
              if (i & 1) continue;
              const int j = i | 1;
              const C x0 = a[i], x1 = a[j];
              if (code == 1) { a[i] = x1; a[j] = x0; }
              else if (code == 2) { a[i] = C(x1.imag(), -x1.real()); a[j] = C(-x0.imag(), x0.real()); }
              else a[j] = -x1;
            }
          }
          for (int k = 0; k < (R.fast == 5 ? 0 : R.ngates); k++) {
            if (R.fast && R.form[k] >= kSlotCxLo) continue;  // bare cx slot: already applied by the permuted load
            const int form = R.fast ? (k == 0 ? 0 : 5) : R.form[k];
            C M[16];
            const double2 *m = p.mats[R.gate[k]];
            for (int i = 0; i < 16; i++) M[i] = C((T)m[i].x, (T)m[i].y);
            if (form >= 10 && form < 14) {
              const int code = codes[(size_t)R.gate[k] * p.nstates + (t >> p.state_shift)], P = form - 10;
              for (int i = 0; i < na && code; i++) {
                if (i & (1 << P)) continue;
                const int j = i | (1 << P);
                const C x0 = a[i], x1 = a[j];
                if (code == 1) { a[i] = x1; a[j] = x0; }
                else if (code == 2) { a[i] = C(x1.imag(), -x1.real()); a[j] = C(-x0.imag(), x0.real()); }
                else a[j] = -x1;
              }
            } else if (form < 6) dense2(pair_of[form][0], pair_of[form][1], M);
            else if (form < 10) {
              const int P = form - 6;
              for (int i = 0; i < na; i++) {
                if (i & (1 << P)) continue;
                const C x0 = a[i], x1 = a[i | (1 << P)];
                a[i] = M[0] * x0 + M[1] * x1;
                a[i | (1 << P)] = M[2] * x0 + M[3] * x1;
              }
            } else if (form < 20) {
              const int P0 = pair_of[form - 14][0], P1 = pair_of[form - 14][1];
              for (int i = 0; i < na; i++) a[i] *= M[((i >> P0) & 1) | (((i >> P1) & 1) << 1)];
            } else {
              const int P = form - 20;
              for (int i = 0; i < na; i++) a[i] *= M[(i >> P) & 1];
            }
          }
        }
        for (int e = 0; e < 16; e++)
          for (int c = 0; c < kPer; c++) tile[(size_t)(base ^ R.eoff[e]) * kPer + c] = a[e * kPer + c];
      }
      if (R.sync) std::fill(owner.begin(), owner.end(), -1);  // CTA barrier: slots may change warps
    }
    // ... and out with the compute-group mapping (thread tid moves slots tid + 256 m): both must be the same bijection
    for (int tid = 0; tid < 256; tid++) {
      uint64_t glo = 0;
      for (int u = 0; u < 8; u++)
        if ((tid >> u) & 1) glo |= p.goff_lo[u];
      const uint32_t slo = phys_slot((uint32_t)tid);
      for (int m = 0; m < kHiCount; m++)
        for (int c = 0; c < kPer; c++)
          psi[((tbase | glo) + p.goff_hi[m]) * kPer + c] = tile[(size_t)(slo ^ p.soff_hi[m]) * kPer + c];
    }
  }
}

// Thread-id bits 0..4 (lane) and 5.. (warp) of a round map to the tile positions that are not round positions.
// `wpos` (kTB - 9 positions, or empty) pins the warp-id bits: consecutive rounds that share wpos keep every warp
// inside its own 2^9-amplitude sub-tile (round positions + lane positions = all positions outside wpos), so only
// __syncwarp() is needed between them.  Returns the padded, sorted round positions.
static std::vector<int> build_round(TileRound &R, const std::vector<int> &round_pos /*tile-local, <=4*/,
                                    const std::vector<int> &wpos, int kTB, bool keep_order = false) {
  auto has = [](const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
  // pad the round to 4 positions with unused tile positions (highest first)
  std::vector<int> pos = round_pos;
  for (int u = kTB - 1; (int)pos.size() < kRoundBits && u >= 0; u--)
    if (!has(pos, u) && !has(wpos, u)) pos.push_back(u);
  if (!keep_order) std::sort(pos.begin(), pos.end());  // fast rounds fix round bit i <-> round_pos[i]
  for (int e = 0; e < 16; e++) {
    uint32_t j = 0;
    for (int i = 0; i < 4; i++)
      if ((e >> i) & 1) j |= 1u << pos[i];
    R.eoff[e] = R.eoff_ld[e] = (uint16_t)phys_slot(j);
  }
  std::vector<int> free_pos;
  for (int u = 0; u < kTB; u++)
    if (!has(pos, u) && !has(wpos, u)) free_pos.push_back(u);
  int lane[3];
  if (!pick_lane_positions(free_pos, lane)) { lane[0] = free_pos[0]; lane[1] = free_pos[1]; lane[2] = free_pos[2]; }
  std::vector<int> tpos(lane, lane + 3);
  for (int u : free_pos)
    if (u != lane[0] && u != lane[1] && u != lane[2]) tpos.push_back(u);
  for (int u : wpos) tpos.push_back(u);
  for (int i = 0; i < kTB - 4; i++) R.gbit[i] = (uint16_t)phys_slot(1u << tpos[i]);
  R.ngates = 0;
  R.sync = 1;
  R.fast = 0;
  R.pre[0] = R.pre[1] = R.pre[2] = R.pre[3] = (uint8_t)kMaxRounds;
  R.pre2[0] = R.pre2[1] = R.pre2[2] = R.pre2[3] = (uint8_t)kMaxRounds;
  R.npre = R.npre2 = 0;
  return pos;
}

static int round_bit_of(const std::vector<int> &sorted_pos, int tile_pos) {
  for (int i = 0; i < (int)sorted_pos.size(); i++)
    if (sorted_pos[i] == tile_pos) return i;
  return -1;
}

// Segments: maximal runs of rounds that leave >= nW = kTB - 9 tile positions untouched; those become the warp-id
// positions of the whole run, which makes the run warp-local (no CTA barrier inside).  round_w[r] = warp positions
// of round r (sorted), seg_end[r] = one past the last round of r's segment.  B200SV_TILE_WARP_LOCAL=0 disables.
static void plan_segments(const std::vector<std::vector<int>> &round_pos, int kTB, std::vector<std::vector<int>> &round_w,
                          std::vector<int> &seg_end) {
  static const int env_wl = [] { const char *e = getenv("B200SV_TILE_WARP_LOCAL"); return e ? atoi(e) : 1; }();
  const int nW = kTB - 9, nr = (int)round_pos.size();
  round_w.assign(nr, {});
  seg_end.assign(nr, 0);
  for (int r0 = 0; r0 < nr;) {
    uint32_t touched = 0;
    int r1 = r0;
    while (r1 < nr) {
      uint32_t t2 = touched;
      for (int u : round_pos[r1]) t2 |= 1u << u;
      if (r1 > r0 && (!env_wl || kTB - __builtin_popcount(t2) < nW)) break;
      touched = t2;
      r1++;
    }
    std::vector<int> cand;
    for (int u = 0; u < kTB; u++)
      if (!((touched >> u) & 1)) cand.push_back(u);
    // choose the warp positions that leave conflict-free lane triples in the most rounds (prefer high positions)
    std::vector<int> best;
    int best_ok = -1;
    const int nc = (int)cand.size();
    for (int a = nc - 1; a >= nW - 1; a--)
      for (int b = (nW >= 2 ? a - 1 : -1); b >= (nW >= 2 ? nW - 2 : -1); b--)
        for (int c = (nW >= 3 ? b - 1 : -1); c >= (nW >= 3 ? 0 : -1); c--) {
          std::vector<int> w;
          w.push_back(cand[a]);
          if (nW >= 2) w.push_back(cand[b]);
          if (nW >= 3) w.push_back(cand[c]);
          int ok = 0;
          for (int r = r0; r < r1; r++) {
            std::vector<int> pos = round_pos[r];
            for (int u = kTB - 1; (int)pos.size() < kRoundBits && u >= 0; u--)
              if (std::find(pos.begin(), pos.end(), u) == pos.end() && std::find(w.begin(), w.end(), u) == w.end())
                pos.push_back(u);
            std::vector<int> fp;
            for (int u = 0; u < kTB; u++)
              if (std::find(pos.begin(), pos.end(), u) == pos.end() && std::find(w.begin(), w.end(), u) == w.end())
                fp.push_back(u);
            int lane[3];
            ok += pick_lane_positions(fp, lane);
          }
          if (ok > best_ok) { best_ok = ok; best = w; }
          if (nW < 3) break;
        }
    std::sort(best.begin(), best.end());
    for (int r = r0; r < r1; r++) { round_w[r] = best; seg_end[r] = r1; }
    r0 = r1;
  }
}

// ------------------------------------------------------------------------------------------ slot rounds
// Round formation for passes without diagonal 2-qubit gates: every gate round is "fast" by construction.  A round has
// two slots, round bits (0,1) and (2,3); a slot holds one dense 2-qubit gate or up to two 1-qubit gates on different
// qubits, which the host multiplies into one 4x4 (v x u: the same 16 DFMA per amplitude as the two gates apart, but no
// form dispatch).  Per-state Pauli ops (sampled noise) get rounds of their own, which cost a byte load unless the state
// drew a non-identity Pauli.  Fills p.rounds / p.mats / p.nrounds; returns true when the pass contains Pauli rounds.
// 0, or kSlotCxLo / kSlotCxHi when g is exactly CX controlled by its first / second qubit (column-major 4x4, index =
// bit(q[0]) + 2 bit(q[1]))
static int bare_cx_form(const QGate &g) {
  if (!g.mat || g.nq != 2) return 0;
  static const int lo[4] = {0, 3, 2, 1}, hi[4] = {0, 1, 3, 2};  // row holding the 1 of column c
  bool is_lo = true, is_hi = true;
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      const double re = g.mat[2 * (r + 4 * c)], im = g.mat[2 * (r + 4 * c) + 1];
      if (im != 0.0 || re != (r == lo[c] ? 1.0 : 0.0)) is_lo = false;
      if (im != 0.0 || re != (r == hi[c] ? 1.0 : 0.0)) is_hi = false;
    }
  return is_lo ? kSlotCxLo : is_hi ? kSlotCxHi : 0;
}
static bool build_slot_rounds(TilePassParams &p, const std::vector<QGate> &gates_in, const std::vector<int> &sel_in,
                              const std::vector<int> &tile_bits, std::vector<int> &leftover) {
  constexpr int kTB = 12;
  // bare cx gates as register renamings: only in passes that carry sampled noise (they run the MODE 4 kernels, the
  // only ones that read the slot forms)
  static const int env_cx = [] { const char *e = getenv("B200SV_TILE_CX_SLOTS"); return e ? atoi(e) : 2; }();
  bool allow_cx = false, any_cx = false;
  for (int gi : sel_in) allow_cx = allow_cx || !gates_in[gi].mat;
  allow_cx = allow_cx && env_cx;
  // Folded cx (B200SV_TILE_CX_SLOTS >= 2, default): a bare cx(a, b) whose next gates on a and on b are 1-qubit gates of
  // this pass -- with at most one sampled Pauli per qubit in between, the noise of the cx -- becomes ONE slot: the cx as
  // the slot's load permutation, the Paulis in between applied through the permuted offsets (TileRound::pre2), then
  // u_b x u_a as the slot's 4x4.  The layer pattern of a noisy circuit (cx layer, 1-qubit layer) then costs one
  // shared-memory round trip per two cx + four 1-qubit gates instead of two.  The macro op takes the cx's place in the
  // order; its constituents only share qubits with each other and with ops that stay behind them.
  struct Macro { int cx, u[2], inner[2], form; };
  std::vector<QGate> gates(gates_in);
  std::vector<int> sel(sel_in);
  std::vector<Macro> macros;
  std::vector<std::array<double, 32>> macro_mats;
  const int first_macro = (int)gates_in.size();
  if (allow_cx && env_cx >= 2) {
    macro_mats.reserve(sel_in.size());
    std::vector<char> gone(sel_in.size(), 0);
    std::vector<int> out;
    for (size_t i = 0; i < sel_in.size(); i++) {
      if (gone[i]) continue;
      const QGate &g = gates_in[sel_in[i]];
      const int form = bare_cx_form(g);
      Macro m{sel_in[i], {-1, -1}, {-1, -1}, form};
      size_t at_u[2] = {0, 0}, at_in[2] = {0, 0};
      bool ok = form != 0;
      for (int x = 0; x < 2 && ok; x++) {  // next ops on q[x]: [one Pauli] then a 1-qubit gate
        ok = false;
        for (size_t j = i + 1; j < sel_in.size(); j++) {
          const QGate &h = gates_in[sel_in[j]];
          if (gone[j] || !((qmask(h) >> g.q[x]) & 1)) continue;
          if (!h.mat && m.inner[x] < 0) { m.inner[x] = sel_in[j]; at_in[x] = j; continue; }
          if (h.mat && h.nq == 1) { m.u[x] = sel_in[j]; at_u[x] = j; ok = true; }
          break;
        }
      }
      if (!ok) { out.push_back(sel_in[i]); continue; }
      macro_mats.emplace_back();
      double *M = macro_mats.back().data();  // column-major 4x4, index = bit(q[0]) + 2 bit(q[1]): u1 x u0
      const cd_t *u0 = reinterpret_cast<const cd_t *>(gates_in[m.u[0]].mat), *u1 = reinterpret_cast<const cd_t *>(gates_in[m.u[1]].mat);
      for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
          const cd_t v = u0[(r & 1) + 2 * (c & 1)] * u1[(r >> 1) + 2 * (c >> 1)];
          M[2 * (r + 4 * c)] = v.real();
          M[2 * (r + 4 * c) + 1] = v.imag();
        }
      QGate mg;
      mg.nq = 2; mg.q[0] = g.q[0]; mg.q[1] = g.q[1]; mg.mat = M; mg.slot = 0;
      gates.push_back(mg);
      out.push_back(first_macro + (int)macros.size());
      macros.push_back(m);
      for (int x = 0; x < 2; x++) {
        gone[at_u[x]] = 1;
        if (m.inner[x] >= 0) gone[at_in[x]] = 1;
      }
    }
    sel.swap(out);
  }
  auto macro_of = [&](int gi) -> const Macro * { return gi >= first_macro ? &macros[gi - first_macro] : nullptr; };
  auto slot_form = [&](const std::vector<int> &sl) { return allow_cx && !sl.empty() ? bare_cx_form(gates[sl[0]]) : 0; };
  auto tile_pos = [&](int q) {
    for (int u = 0; u < kTB; u++)
      if (tile_bits[u] == q) return u;
    throw Error("tile pass: qubit not in tile");
  };
  struct SRound { int kind = 0; std::vector<int> slot[2]; int pauli = -1; std::vector<std::pair<int, int>> pre; };
  std::vector<SRound> rounds;
  std::vector<std::vector<int>> round_pos;
  std::vector<int> rem(sel.begin(), sel.end());
  // a per-state Pauli whose next op on its qubit (inside this pass) is a dense gate is folded into that gate's round
  // as a "pre-Pauli" instead of getting a round of its own
  static const int env_fold = [] { const char *e = getenv("B200SV_TILE_FOLD_PAULI"); return e ? atoi(e) : 1; }();
  std::vector<char> attachable(gates.size(), 0);
  {
    int next_dense[64];
    std::fill(next_dense, next_dense + 64, 0);
    for (int k = (int)sel.size() - 1; k >= 0; k--) {
      const QGate &g = gates[sel[k]];
      if (!g.mat) attachable[sel[k]] = env_fold && next_dense[g.q[0]];
      for (int x = 0; x < g.nq; x++) next_dense[g.q[x]] = g.mat != nullptr;
    }
  }
  int nmat = 0, npaul = 0;
  while (!rem.empty()) {
    if ((int)rounds.size() >= kMaxRounds || nmat + 2 > kMaxTileGates || npaul + 8 > kMaxRounds) {
      for (int gi : rem) {  // macro ops go back as their constituents
        const Macro *mc = macro_of(gi);
        if (!mc) { leftover.push_back(gi); continue; }
        leftover.push_back(mc->cx);
        for (int x = 0; x < 2; x++) {
          if (mc->inner[x] >= 0) leftover.push_back(mc->inner[x]);
          leftover.push_back(mc->u[x]);
        }
      }
      break;
    }
    SRound sr;
    uint64_t rq = 0, blocked = 0;
    std::vector<std::pair<int, int>> rest;  // (scan position, op): held Paulis that stay unattached re-enter in order
    int held[64];
    std::fill(held, held + 64, -1);
    std::vector<std::pair<int, int>> held_at;  // (scan position, op)
    int pos = 0;
    for (int gi : rem) {
      const QGate &g = gates[gi];
      const uint64_t m = qmask(g);
      bool taken = false;
      if (!(m & blocked)) {
        if (!g.mat) {
          const int x = g.q[0];
          if (attachable[gi] && sr.kind != 5 && held[x] < 0 && !(m & rq)) {  // wait for the gate on x in this round
            held[x] = gi;
            held_at.push_back({pos, gi});
            taken = true;
          } else if (sr.kind == 0 && held_at.empty()) {  // stand-alone Pauli round
            sr.kind = 5; sr.pauli = gi; rq |= m; taken = true;
          }
        } else if (sr.kind != 5 && !(m & rq)) {
          for (int k = 0; k < 2 && !taken; k++) {
            std::vector<int> &sl = sr.slot[k];
            const bool fits = g.nq == 2 ? sl.empty() : (sl.size() < 2 && (sl.empty() || gates[sl[0]].nq == 1));
            if (fits) { sl.push_back(gi); sr.kind = 1; rq |= m; taken = true; }
          }
          if (taken)
            for (int x = 0; x < g.nq; x++)
              if (held[g.q[x]] >= 0) { sr.pre.push_back({g.q[x], held[g.q[x]]}); held[g.q[x]] = -2; }
        }
      }
      if (!taken) { blocked |= m; rest.push_back({pos, gi}); }
      pos++;
    }
    for (auto &h : held_at)  // Paulis whose gate did not make it into this round
      if (held[gates[h.second].q[0]] == h.second) rest.push_back(h);
    std::sort(rest.begin(), rest.end());
    if (sr.kind == 0) {  // nothing but held Paulis: give the first one a round of its own (progress)
      sr.kind = 5; sr.pauli = rest.front().second; rq |= qmask(gates[sr.pauli]);
      rest.erase(rest.begin());
    }
    if (sr.kind == 1 && sr.slot[0].empty()) std::swap(sr.slot[0], sr.slot[1]);
    if (sr.kind == 1 && !sr.slot[1].empty()) sr.kind = 2;
    std::vector<int> rp;
    for (int q = 0; q < 64; q++)
      if ((rq >> q) & 1) rp.push_back(tile_pos(q));
    if (sr.kind != 5)
      for (int k = 0; k < sr.kind; k++) nmat += slot_form(sr.slot[k]) ? 0 : 1;
    npaul += sr.kind == 5 ? 1 : (int)sr.pre.size();
    if (sr.kind != 5)
      for (int k = 0; k < sr.kind; k++)
        if (const Macro *mc = macro_of(sr.slot[k][0])) npaul += (mc->inner[0] >= 0) + (mc->inner[1] >= 0);
    rounds.push_back(sr);
    round_pos.push_back(rp);
    rem.clear();
    for (auto &e : rest) rem.push_back(e.second);
  }
  const int nr = (int)rounds.size();
  p.nrounds = nr;
  std::vector<std::vector<int>> round_w;
  std::vector<int> seg_end;
  plan_segments(round_pos, kTB, round_w, seg_end);
  auto has = [](const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
  int nm = 0;
  bool any_pauli = false;
  p.npauli = 0;
  std::fill(p.pauli_slot, p.pauli_slot + kMaxRounds, (uint16_t)0);
  for (int r = 0; r < nr; r++) {
    const SRound &sr = rounds[r];
    std::vector<int> pads;  // unused positions outside the warp positions, highest first
    for (int u = kTB - 1; u >= 0; u--)
      if (!has(round_pos[r], u) && !has(round_w[r], u)) pads.push_back(u);
    size_t np = 0;
    std::vector<int> ordered;
    TileRound &R = p.rounds[r];
    if (sr.kind == 5) {
      ordered.push_back(tile_pos(gates[sr.pauli].q[0]));
      while (ordered.size() < 4) ordered.push_back(pads[np++]);
      build_round(R, ordered, round_w[r], kTB, true);
      R.fast = 5;
      if (p.npauli >= kMaxRounds) throw Error("tile pass: Pauli table overflow");
      p.pauli_slot[p.npauli] = (uint16_t)gates[sr.pauli].slot;
      R.gate[0] = (uint16_t)p.npauli++;
      R.ngates = 0;
      any_pauli = true;
    } else {
      cd_t M4[2][16];
      for (int k = 0; k < sr.kind; k++) {
        const std::vector<int> &sl = sr.slot[k];
        const QGate &g0 = gates[sl[0]];
        if (g0.nq == 2) {
          ordered.push_back(tile_pos(g0.q[0]));
          ordered.push_back(tile_pos(g0.q[1]));
          std::copy(reinterpret_cast<const cd_t *>(g0.mat), reinterpret_cast<const cd_t *>(g0.mat) + 16, M4[k]);
        } else {
          const cd_t *u = reinterpret_cast<const cd_t *>(g0.mat);
          static const cd_t I2[4] = {1, 0, 0, 1};
          const cd_t *v = sl.size() == 2 ? reinterpret_cast<const cd_t *>(gates[sl[1]].mat) : I2;
          ordered.push_back(tile_pos(g0.q[0]));
          ordered.push_back(sl.size() == 2 ? tile_pos(gates[sl[1]].q[0]) : pads[np++]);
          for (int c = 0; c < 4; c++)      // column-major 4x4, index = bit(first qubit) + 2 bit(second): v x u
            for (int rr = 0; rr < 4; rr++) M4[k][rr + 4 * c] = u[(rr & 1) + 2 * (c & 1)] * v[(rr >> 1) + 2 * (c >> 1)];
        }
      }
      while (ordered.size() < 4) ordered.push_back(pads[np++]);
      build_round(R, ordered, round_w[r], kTB, true);
      R.fast = (uint8_t)sr.kind;
      R.ngates = (uint8_t)sr.kind;
      for (const auto &pr : sr.pre) {  // (qubit, Pauli op): applied to that qubit's round bit before the gates
        const int bit = (int)(std::find(ordered.begin(), ordered.end(), tile_pos(pr.first)) - ordered.begin());
        if (p.npauli >= kMaxRounds) throw Error("tile pass: Pauli table overflow");
        p.pauli_slot[p.npauli] = (uint16_t)gates[pr.second].slot;
        R.pre[bit] = (uint8_t)p.npauli++;
        R.npre++;
        any_pauli = true;
      }
      for (int k = 0; k < sr.kind; k++) {
        if (const int cxf = slot_form(sr.slot[k])) {  // no matrix, no arithmetic: output e = input e ^ (ctl(e) << tgt)
          R.form[k] = (uint8_t)cxf;
          R.gate[k] = 0;
          any_cx = true;
          const int ctl = 2 * k + (cxf == kSlotCxLo ? 0 : 1), tgt = 2 * k + (cxf == kSlotCxLo ? 1 : 0);
          uint16_t ld[16];
          for (int e = 0; e < 16; e++) ld[e] = R.eoff_ld[e ^ (((e >> ctl) & 1) << tgt)];
          std::copy(ld, ld + 16, R.eoff_ld);
          continue;
        }
        if (nm >= kMaxTileGates) throw Error("tile pass: matrix slots exhausted");
        double2 *M = p.mats[nm];
        for (int i = 0; i < 4; i++)
          for (int j = 0; j < 4; j++) M[i * 4 + j] = mk<double>(M4[k][i + 4 * j].real(), M4[k][i + 4 * j].imag());
        R.form[k] = (uint8_t)(k == 0 ? 0 : 5);
        R.gate[k] = (uint16_t)nm++;
        if (const Macro *mc = macro_of(sr.slot[k][0])) {  // folded cx: load permutation, then the noise in between
          const int ctl = 2 * k + (mc->form == kSlotCxLo ? 0 : 1), tgt = 2 * k + (mc->form == kSlotCxLo ? 1 : 0);
          uint16_t ld[16];
          for (int e = 0; e < 16; e++) ld[e] = R.eoff_ld[e ^ (((e >> ctl) & 1) << tgt)];
          std::copy(ld, ld + 16, R.eoff_ld);
          for (int x = 0; x < 2; x++) {
            if (mc->inner[x] < 0) continue;
            if (p.npauli >= kMaxRounds) throw Error("tile pass: Pauli table overflow");
            p.pauli_slot[p.npauli] = (uint16_t)gates[mc->inner[x]].slot;
            R.pre2[2 * k + x] = (uint8_t)p.npauli++;
            R.npre2++;
            any_pauli = true;
          }
          any_cx = true;
        }
      }
    }
    R.sync = (r == seg_end[r] - 1);
  }
  return any_pauli || any_cx;
}

// ------------------------------------------------------------------------------------------ planned passes
// A pass is fully described by its parameter block and the kernel variant that consumes it.  Passes are launched at
// once (the C ABI's immediate path) or captured into a TilePlan (State::capture) and launched later, possibly several
// times and possibly restricted to a SLAB of the state: a sub-cube in which some global index bits outside the tile are
// held fixed.  A slab launch only changes how tile indices map to addresses -- the fixed positions join the tile's
// zero-insertion list, the fixed bit values become a pointer offset, the tile count shrinks -- so the kernels are
// unchanged.  Slabs are what lets the sharded executor overlap a global-qubit exchange with the passes next to it.
enum TileVariant { TV_PIPE2_PAULI, TV_PIPE_PAULI, TV_PIPE2_FAST, TV_PIPE_FAST, TV_PIPE_GENERIC, TV_PASS12, TV_PASS11,
                   TV_PIPE2_F32, TV_DENSE, TV_PIPE3_FAST };
struct TilePlannedPass {
  TileVariant variant;
  TilePassParams p;        // unused for TV_DENSE
  int dq[2] = {0, 0}, dnq = 0;
  double dmat[32];         // TV_DENSE: one streaming pass for one gate (states too small for tiles)
};
struct TilePlan {
  std::vector<TilePlannedPass> passes;
};

static void launch_tile_variant(State &s, const TilePassParams &p, TileVariant v, double2 *psi, int grid_cap) {
  NvtxRange nvtx("b200sv tile pass");
  // function attributes are per device: one flag per device ordinal (a process may drive several GPUs)
  static bool attr_set_dev[64] = {};
  bool &attr_set = attr_set_dev[s.device & 63];
  const int smem_pipe = kPipeBufs * (16 << 12) + 64;
  const int smem_pauli = smem_pipe + 4 * (kMaxRounds + 16);
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(tile_pass_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (16 << 12)));
    B200_CUDA(cudaFuncSetAttribute(tile_pass_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (16 << 11)));
    B200_CUDA(cudaFuncSetAttribute(tile_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pipe));
    B200_CUDA(cudaFuncSetAttribute(tile_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pipe));
    B200_CUDA(cudaFuncSetAttribute(tile_pipe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pauli));
    B200_CUDA(cudaFuncSetAttribute(tile_pipe2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pauli));
    B200_CUDA(cudaFuncSetAttribute(tile_pipe2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pipe));
    B200_CUDA(cudaFuncSetAttribute(tile_pipe2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pipe));
    B200_CUDA(cudaFuncSetAttribute(tile_pipe2_kernel<1, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * (16 << 11) + 128));
    attr_set = true;
  }
  const uint64_t sms = (uint64_t)std::max(1, grid_cap > 0 ? std::min(grid_cap, s.num_sms) : s.num_sms);
  const int grid = (int)std::min<uint64_t>(p.ntiles, sms);
  switch (v) {
  case TV_PIPE2_PAULI: tile_pipe2_kernel<4><<<grid, 640, smem_pauli, s.stream>>>(psi, p); break;
  case TV_PIPE_PAULI: tile_pipe_kernel<4><<<grid, 512, smem_pauli, s.stream>>>(psi, p); break;
  case TV_PIPE2_FAST: tile_pipe2_kernel<1><<<grid, 640, smem_pipe, s.stream>>>(psi, p); break;
  case TV_PIPE_FAST: tile_pipe_kernel<1><<<grid, 512, smem_pipe, s.stream>>>(psi, p); break;
  case TV_PIPE_GENERIC: tile_pipe_kernel<2><<<grid, 512, smem_pipe, s.stream>>>(psi, p); break;
  case TV_PIPE2_F32: tile_pipe2_kernel<3><<<grid, 640, smem_pipe, s.stream>>>(psi, p); break;
  case TV_PIPE3_FAST: tile_pipe2_kernel<1, 11><<<grid, 640, 6 * (16 << 11) + 128, s.stream>>>(psi, p); break;
  case TV_PASS12:
    tile_pass_kernel<12><<<(int)std::min<uint64_t>(p.ntiles, sms * 2), 256, 16 << 12, s.stream>>>(psi, p);
    break;
  case TV_PASS11:
    tile_pass_kernel<11><<<(int)std::min<uint64_t>(p.ntiles, sms * 4), 128, 16 << 11, s.stream>>>(psi, p);
    break;
  default: throw Error("tile pass: bad kernel variant");
  }
  B200_CUDA(cudaGetLastError());
}

// launch now, or record into the plan being captured
static void emit_tile_pass(State &s, const TilePassParams &p, TileVariant v) {
  if (s.capture) {
    s.capture->passes.emplace_back();
    s.capture->passes.back().variant = v;
    s.capture->passes.back().p = p;
    return;
  }
  launch_tile_variant(s, p, v, (double2 *)s.data, 0);
}

// One pass: gates[sel] all fit the tile; tile_bits sorted global positions (kTB of them).
// Returns the entries of `sel` that did not fit (round or matrix-slot budget): the caller re-queues them.
static std::vector<int> run_tile_pass(State &s, const std::vector<QGate> &gates, const std::vector<int> &sel,
                          const std::vector<int> &tile_bits, const uint8_t *dev_codes, int kTB) {
  const int kLoBits = kTB - 4;
  static thread_local TilePassParams p;
  p.ntiles = s.total_amps() >> kTB;
  p.codes = dev_codes;
  p.nstates = (uint64_t)s.nstates;
  p.state_shift = s.nq - kTB;
  p.ins.n = kTB;
  for (int u = 0; u < kTB; u++) p.ins.pos[u] = (uint8_t)tile_bits[u];
  for (int u = 0; u < kLoBits; u++) p.goff_lo[u] = 1ull << tile_bits[u];
  for (int m = 0; m < kHiCount; m++) {
    uint64_t go = 0;
    for (int b = 0; b < kTB - kLoBits; b++)
      if ((m >> b) & 1) go |= 1ull << tile_bits[kLoBits + b];
    p.goff_hi[m] = go;
    p.soff_hi[m] = (uint16_t)phys_slot((uint32_t)m << kLoBits);
  }
  auto tile_pos = [&](int q) {
    for (int u = 0; u < kTB; u++)
      if (tile_bits[u] == q) return u;
    throw Error("tile pass: qubit not in tile");
  };
  static const int env_pipe = [] { const char *e = getenv("B200SV_TILE_PIPE"); return e ? atoi(e) : 2; }();
  static const int env_slots = [] { const char *e = getenv("B200SV_TILE_SLOTS"); return e ? atoi(e) : 1; }();
  bool use_slots = env_slots && kTB == 12 && env_pipe != 0, slot_pauli = false;
  for (int gi : sel)
    if (gates[gi].mat && gates[gi].nq == 2 && is_diag(gates[gi])) use_slots = false;  // cz / cp / rzz: one-multiply forms
  std::vector<int> leftover;
  if (use_slots) slot_pauli = build_slot_rounds(p, gates, sel, tile_bits, leftover);
  int ndense = 0;
  // rounds: scan in order, capacity 4 tile positions, respect dependencies
  std::vector<int> rem(sel.size());
  for (size_t i = 0; i < sel.size(); i++) rem[i] = (int)i;  // indices into sel
  if (use_slots) rem.clear();
  std::vector<std::vector<int>> round_take, round_pos;
  // A round that starts with a dense (non-diagonal) 2-qubit gate stays "fast": at most one more such gate on two
  // other qubits joins it (straight-line code, no form dispatch, no register shuffling at the dispatch joins);
  // further gates on the same qubits go to the next round, which costs one warp-local shared-memory round trip --
  // cheaper than the generic dispatch.  Rounds that start with a 1-qubit / diagonal / Pauli op are generic.
  static const int env_prefer_fast = [] { const char *e = getenv("B200SV_TILE_PREFER_FAST"); return e ? atoi(e) : 1; }();
  auto fast_eligible = [&](const QGate &g) { return g.mat && g.nq == 2 && !is_diag(g); };
  while (!rem.empty()) {
    if ((int)round_take.size() >= kMaxRounds) {  // out of rounds: hand the rest back
      for (int li : rem) leftover.push_back(sel[li]);
      break;
    }
    uint64_t rq = 0, blocked = 0;
    std::vector<int> take, rest;
    bool fast_round = false;
    for (int li : rem) {
      const QGate &g = gates[sel[li]];
      const uint64_t m = qmask(g);
      if ((m & blocked) || (int)take.size() >= kMaxRoundGates) { blocked |= m; rest.push_back(li); continue; }
      if (take.empty()) fast_round = env_prefer_fast && fast_eligible(g);
      else if (fast_round && (!fast_eligible(g) || (m & rq) || take.size() >= 2)) { blocked |= m; rest.push_back(li); continue; }
      if (__builtin_popcountll(rq | m) <= kRoundBits) { rq |= m; take.push_back(li); }
      else { blocked |= m; rest.push_back(li); }
    }
    std::vector<int> rpos;
    for (int q = 0; q < 64; q++)
      if ((rq >> q) & 1) rpos.push_back(tile_pos(q));
    round_take.push_back(take);
    round_pos.push_back(rpos);
    rem.swap(rest);
  }
  // segments of warp-local rounds and their warp-id positions
  const int nr = (int)round_take.size();
  if (!use_slots) p.nrounds = nr;
  std::vector<std::vector<int>> round_w;
  std::vector<int> seg_end;
  plan_segments(round_pos, kTB, round_w, seg_end);
  for (int r0 = 0; r0 < nr;) {
    const int r1 = seg_end[r0];
    const std::vector<int> &best = round_w[r0];
    for (int r = r0; r < r1; r++) {
      TileRound &R = p.rounds[r];
      // fast round: one or two disjoint dense 2-qubit gates -> round bits (0,1) and (2,3) in gate-qubit order
      static const int env_fast = [] { const char *e = getenv("B200SV_TILE_FAST"); return e ? atoi(e) : 1; }();
      int fast = 0;
      std::vector<int> ordered;
      if (env_fast && round_take[r].size() <= 2) {
        fast = (int)round_take[r].size();
        for (int li : round_take[r]) {
          const QGate &g = gates[sel[li]];
          if (!g.mat || g.nq != 2 || is_diag(g)) { fast = 0; break; }
          ordered.push_back(tile_pos(g.q[0]));
          ordered.push_back(tile_pos(g.q[1]));
        }
        if (fast == 2 && (ordered[0] == ordered[2] || ordered[0] == ordered[3] || ordered[1] == ordered[2] ||
                          ordered[1] == ordered[3]))
          fast = 0;
      }
      const std::vector<int> pos = build_round(R, fast ? ordered : round_pos[r], best, kTB, fast != 0);
      R.sync = (r == r1 - 1);
      R.fast = (uint8_t)fast;
      for (int li : round_take[r]) {
        const QGate &g = gates[sel[li]];
        if (!g.mat) {  // per-state Pauli
          R.form[R.ngates] = (uint8_t)(10 + round_bit_of(pos, tile_pos(g.q[0])));
          R.gate[R.ngates++] = (uint16_t)g.slot;
          continue;
        }
        if (ndense >= kMaxTileGates) throw Error("tile pass: too many dense gates");
        const int mi = ndense++;
        double2 *M = p.mats[mi];
        const bool diag = is_diag(g);
        if (diag && g.nq == 1) {
          const int b = round_bit_of(pos, tile_pos(g.q[0]));
          for (int i = 0; i < 2; i++) M[i] = mk<double>(g.mat[2 * (i + 2 * i)], g.mat[2 * (i + 2 * i) + 1]);
          R.form[R.ngates] = (uint8_t)(20 + b);
        } else if (diag) {
          int b0 = round_bit_of(pos, tile_pos(g.q[0])), b1 = round_bit_of(pos, tile_pos(g.q[1]));
          const bool flip = b0 > b1;  // canonical: diagonal index bit0 <-> lower round bit
          for (int i = 0; i < 4; i++) {
            const int si = flip ? ((i >> 1) | ((i & 1) << 1)) : i;
            M[i] = mk<double>(g.mat[2 * (si + 4 * si)], g.mat[2 * (si + 4 * si) + 1]);
          }
          if (flip) std::swap(b0, b1);
          static const int form_of[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
          R.form[R.ngates] = (uint8_t)(14 + form_of[b0][b1]);
        } else if (g.nq == 1) {
          const int b = round_bit_of(pos, tile_pos(g.q[0]));
          for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) M[i * 2 + j] = mk<double>(g.mat[2 * (i + 2 * j)], g.mat[2 * (i + 2 * j) + 1]);
          R.form[R.ngates] = (uint8_t)(6 + b);
        } else {
          int b0 = round_bit_of(pos, tile_pos(g.q[0])), b1 = round_bit_of(pos, tile_pos(g.q[1]));
          const bool flip = b0 > b1;  // canonical: matrix bit0 <-> lower round bit
          for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
              const int si = flip ? ((i >> 1) | ((i & 1) << 1)) : i, sj = flip ? ((j >> 1) | ((j & 1) << 1)) : j;
              M[i * 4 + j] = mk<double>(g.mat[2 * (si + 4 * sj)], g.mat[2 * (si + 4 * sj) + 1]);
            }
          if (flip) std::swap(b0, b1);
          static const int form_of[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
          R.form[R.ngates] = (uint8_t)form_of[b0][b1];
        }
        R.gate[R.ngates++] = (uint16_t)mi;
      }
    }
    r0 = r1;
  }
  static const int env_trace_rounds = [] { const char *e = getenv("B200SV_TILE_TRACE"); return e ? atoi(e) : 0; }();
  if (env_trace_rounds >= 2) {  // round census: dense slots / bare-cx slots / stand-alone Pauli rounds / folded Paulis / barriers
    int dense = 0, cx = 0, prounds = 0, pre = 0, syncs = 0, cxonly = 0;
    for (int r = 0; r < p.nrounds; r++) {
      const TileRound &R = p.rounds[r];
      syncs += R.sync;
      pre += R.npre + R.npre2;
      if (R.fast == 5) { prounds++; continue; }
      int d = 0, c = 0;
      for (int k = 0; k < R.ngates; k++) (R.fast && slot_pauli && R.form[k] >= kSlotCxLo ? c : d)++;
      dense += d; cx += c;
      cxonly += (d == 0 && c > 0);
    }
    fprintf(stderr, "b200sv   rounds %d: dense slots %d, cx slots %d (%d cx-only rounds), Pauli rounds %d, folded Paulis %d, CTA barriers %d\n",
            p.nrounds, dense, cx, cxonly, prounds, pre, syncs);
  }
  if (s.selftest_host) {  // scheduler self-test: interpret the parameter block on the host array
    if (!s.plan_only) emulate_tile_pass<double>(p, s.selftest_host, s.selftest_codes, false);
    return leftover;
  }
  bool all_fast = true;
  for (int r = 0; r < p.nrounds; r++) all_fast = all_fast && p.rounds[r].fast;
  TileVariant v;
  if (slot_pauli) v = env_pipe == 2 ? TV_PIPE2_PAULI : TV_PIPE_PAULI;   // fast rounds + sampled-noise Paulis
  else if (kTB == 12 && env_pipe == 2 && all_fast) v = TV_PIPE2_FAST;   // memory-warp variant: fast-only code fits its register budget
  else if (kTB == 12 && env_pipe) v = all_fast ? TV_PIPE_FAST : TV_PIPE_GENERIC;
  else if (kTB == 11 && env_pipe == 3 && all_fast) v = TV_PIPE3_FAST;
  else v = kTB == 12 ? TV_PASS12 : TV_PASS11;
  emit_tile_pass(s, p, v);
  return leftover;
}

// ------------------------------------------------------------------------------------------ pass packing
// Which ops ride on the next pass?  Ready-set greedy over the dependency DAG (ops ordered through shared qubits):
// repeatedly add the ready op that brings the fewest NEW qubits into the tile (ties: the one with most successors that
// would already fit, then program order), so a tile first deepens -- later-layer gates on qubits it already holds are
// free -- before it widens.  Against first-fit in program order this saves a quarter of the passes of a Quantum Volume
// circuit (10 seeds, 33 qubits: 23.3 -> 17.2 passes per circuit).  The pick order is a topological order, which is
// what the round formation needs.  B200SV_TILE_PACK=0 restores first-fit.
struct PassPacker {
  const std::vector<QGate> &gates;
  std::vector<std::vector<int>> succ;
  std::vector<int> pending;  // undone predecessors
  std::vector<char> done;
  int ndone = 0;
  bool noisy = false;  // the sequence carries sampled-noise Paulis: slot rounds with bare-cx slots (build_slot_rounds)
  // matrix-slot cost of an op in half slots: two 1-qubit gates share a 4x4, a bare cx of a noisy pass needs none
  int half_slots(const QGate &g) const {
    if (!g.mat) return 0;
    if (!noisy) return 2;
    return g.nq == 1 ? 1 : (bare_cx_form(g) ? 0 : 2);
  }
  explicit PassPacker(const std::vector<QGate> &g) : gates(g), succ(g.size()), pending(g.size(), 0), done(g.size(), 0) {
    static const int env_cx = [] { const char *e = getenv("B200SV_TILE_CX_SLOTS"); return e ? atoi(e) : 1; }();
    bool any_diag2 = false;  // diagonal 2-qubit gates send their pass to the generic rounds: one matrix per gate there
    for (const QGate &x : g) {
      noisy = noisy || (!x.mat && env_cx);
      any_diag2 = any_diag2 || (x.mat && x.nq == 2 && is_diag(x));
    }
    static const bool slot_rounds = [] {  // the debugging knobs that turn slot rounds off (run_tile_pass)
      const char *a = getenv("B200SV_TILE_SLOTS"), *b = getenv("B200SV_TILE_PIPE");
      return (a ? atoi(a) : 1) != 0 && (b ? atoi(b) : 2) != 0;
    }();
    noisy = noisy && !any_diag2 && slot_rounds;
    int last[64];
    std::fill(last, last + 64, -1);
    for (int i = 0; i < (int)g.size(); i++)
      for (int x = 0; x < g[i].nq; x++) {
        const int q = g[i].q[x];
        if (last[q] >= 0 && (x == 0 || last[q] != last[g[i].q[0]])) { succ[last[q]].push_back(i); pending[i]++; }
        last[q] = i;
      }
  }
  bool finished() const { return ndone == (int)gates.size(); }
  // selects ops for one pass; Q (in/out) = tile qubit mask, cap = tile bits
  std::vector<int> select(uint64_t &Q, int cap, int max_dense, int max_ops) {
    // Noisy sequences: a sampled Pauli whose next op on its qubit does not make it into this pass would get a round of
    // its own at the end of the pass; left for the next pass it is folded into that op's round.  Such Paulis are banned
    // and the selection is redone, so that the op slots they held go to gates (a ban never unblocks anything: the ops
    // behind a banned Pauli were not selectable in the first place).
    std::vector<char> banned(gates.size(), 0);
    const uint64_t Q0 = Q;
    std::vector<int> sel;
    for (int attempt = 0; attempt < 6; attempt++) {
      Q = Q0;
      sel = select_once(Q, cap, max_dense, max_ops, banned);
      static const int env_defer = [] { const char *e = getenv("B200SV_TILE_DEFER_PAULI"); return e ? atoi(e) : 1; }();
      if (!noisy || !env_defer) break;
      std::vector<char> in(gates.size(), 0);
      for (int i : sel) in[i] = 1;
      int ndrop = 0, ngate = 0;
      std::vector<int> drop;
      for (int k = (int)sel.size() - 1; k >= 0; k--) {  // reverse: a dropped Pauli exposes the Pauli before it
        const int i = sel[k];
        bool d = false;
        if (!gates[i].mat)
          for (int sx : succ[i]) d = d || !in[sx];
        if (d) { in[i] = 0; drop.push_back(i); ndrop++; }
        else ngate += gates[i].mat != nullptr;
      }
      if (!ndrop || !ngate) break;  // nothing to defer, or a pass of Paulis only: run it as selected
      for (int i : drop) banned[i] = 1;
    }
    return sel;
  }
  std::vector<int> select_once(uint64_t &Q, int cap, int max_dense, int max_ops, const std::vector<char> &banned) {
    std::vector<int> pp(pending), cand, sel;
    for (int i = 0; i < (int)gates.size(); i++)
      if (!done[i] && pp[i] == 0 && !banned[i]) cand.push_back(i);
    int ndense = 0;
    while ((int)sel.size() < max_ops) {
      int best = -1, best_new = 99, best_sc = -1;
      size_t best_at = 0;
      for (size_t c = 0; c < cand.size(); c++) {
        const int i = cand[c];
        const uint64_t m = qmask(gates[i]);
        if (gates[i].mat && ndense + half_slots(gates[i]) > 2 * max_dense) continue;
        if (__builtin_popcountll(Q | m) > cap) continue;
        const int nw = __builtin_popcountll(m & ~Q);
        int sc = 0;
        for (int sx : succ[i])
          if (!(qmask(gates[sx]) & ~(Q | m))) sc++;
        if (nw < best_new || (nw == best_new && (sc > best_sc || (sc == best_sc && i < best)))) {
          best = i; best_new = nw; best_sc = sc; best_at = c;
        }
      }
      if (best < 0) break;
      cand.erase(cand.begin() + best_at);
      sel.push_back(best);
      Q |= qmask(gates[best]);
      ndense += half_slots(gates[best]);
      for (int sx : succ[best])
        if (--pp[sx] == 0 && !banned[sx]) cand.push_back(sx);
    }
    return sel;
  }
  void commit(const std::vector<int> &sel, const std::vector<int> &not_run) {
    for (int i : sel) {
      if (std::find(not_run.begin(), not_run.end(), i) != not_run.end()) continue;
      done[i] = 1;
      ndone++;
      for (int sx : succ[i]) pending[sx]--;
    }
  }
};

// ------------------------------------------------------------------------------------------ single-precision passes
// One pass over a float state: tile_bits = 13 sorted global positions, tile_bits[0] == 0 (global qubit 0 lives inside
// the 16-byte slot).  Slot space: slot bit u <-> tile_bits[u + 1].  Returns the gates that did not fit.
static std::vector<int> run_tile_pass_f32(State &s, const std::vector<QGate> &gates, const std::vector<int> &sel,
                                          const std::vector<int> &tile_bits) {
  constexpr int kSB = 12;  // slot bits
  static thread_local TilePassParams p;
  p.ntiles = s.total_amps() >> (kSB + 1);
  p.codes = nullptr;
  p.nstates = (uint64_t)s.nstates;
  p.state_shift = s.nq - (kSB + 1);
  p.ins.n = kSB;
  for (int u = 0; u < kSB; u++) p.ins.pos[u] = (uint8_t)(tile_bits[u + 1] - 1);
  for (int u = 0; u < 8; u++) p.goff_lo[u] = 1ull << (tile_bits[u + 1] - 1);
  for (int m = 0; m < kHiCount; m++) {
    uint64_t go = 0;
    for (int b = 0; b < 4; b++)
      if ((m >> b) & 1) go |= 1ull << (tile_bits[8 + b + 1] - 1);
    p.goff_hi[m] = go;
    p.soff_hi[m] = (uint16_t)phys_slot((uint32_t)m << 8);
  }
  auto slot_pos = [&](int q) {  // -1: global qubit 0 (inside the slot)
    for (int u = 0; u <= kSB; u++)
      if (tile_bits[u] == q) return u - 1;
    throw Error("tile pass: qubit not in tile");
  };
  // rounds: gate A (any), optionally gate B on two other slot positions (B never involves qubit 0)
  struct FRound { int a = -1, b = -1; };
  std::vector<FRound> frounds;
  std::vector<std::vector<int>> round_pos;
  std::vector<int> rem(sel.begin(), sel.end()), leftover;
  int ndense = 0;
  while (!rem.empty()) {
    if ((int)frounds.size() >= kMaxRounds || ndense + 2 > kMaxTileGates) { leftover = rem; break; }
    FRound fr;
    uint64_t rq = 0, blocked = 0;
    std::vector<int> rest;
    for (int gi : rem) {
      const QGate &g = gates[gi];
      const uint64_t m = qmask(g);
      if (m & blocked) { blocked |= m; rest.push_back(gi); continue; }
      if (fr.a < 0) { fr.a = gi; rq |= m; }
      else if (fr.b < 0 && !(m & 1ull) && !(m & rq)) { fr.b = gi; rq |= m; }
      else { blocked |= m; rest.push_back(gi); }
    }
    std::vector<int> rp;
    for (int q = 1; q < 64; q++)
      if ((rq >> q) & 1) rp.push_back(slot_pos(q));
    frounds.push_back(fr);
    round_pos.push_back(rp);
    ndense += 1 + (fr.b >= 0);
    rem.swap(rest);
  }
  const int nr = (int)frounds.size();
  p.nrounds = nr;
  std::vector<std::vector<int>> round_w;
  std::vector<int> seg_end;
  plan_segments(round_pos, kSB, round_w, seg_end);
  int nm = 0;
  auto put_matrix = [&](const QGate &g, bool q0_is_low_round_bit) {
    // 4x4 row-major, matrix bit 0 <-> lower round bit; 1-qubit gates become (identity on the pad bit) x U; every entry
    // as the two FFMA2 multiplier pairs {re, re, -im, im} (apply2f)
    float4 *M = reinterpret_cast<float4 *>(p.mats[nm]);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        double re, im;
        if (g.nq == 1) {
          const bool same_hi = (i >> 1) == (j >> 1);
          re = same_hi ? g.mat[2 * ((i & 1) + 2 * (j & 1))] : 0.0;
          im = same_hi ? g.mat[2 * ((i & 1) + 2 * (j & 1)) + 1] : 0.0;
        } else {
          const int si = q0_is_low_round_bit ? i : ((i >> 1) | ((i & 1) << 1));
          const int sj = q0_is_low_round_bit ? j : ((j >> 1) | ((j & 1) << 1));
          re = g.mat[2 * (si + 4 * sj)];
          im = g.mat[2 * (si + 4 * sj) + 1];
        }
        M[i * 4 + j] = make_float4((float)re, (float)re, -(float)im, (float)im);
      }
    return nm++;
  };
  for (int r = 0; r < nr; r++) {
    const FRound &fr = frounds[r];
    const QGate &A = gates[fr.a];
    const std::vector<int> &w = round_w[r];
    auto has = [](const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    std::vector<int> pads;  // unused slot positions outside the warp positions, highest first
    for (int u = kSB - 1; u >= 0; u--)
      if (!has(round_pos[r], u) && !has(w, u)) pads.push_back(u);
    size_t np = 0;
    std::vector<int> ordered(4);
    TileRound &R = p.rounds[r];
    const bool a_has0 = (qmask(A) & 1ull) != 0;
    bool a_q0_low;
    if (a_has0) {  // bits (0,1): qubit 0 and the other qubit (or a pad for a 1-qubit gate on qubit 0)
      if (A.nq == 1) { ordered[0] = pads[np++]; a_q0_low = true; }
      else { ordered[0] = slot_pos(A.q[0] == 0 ? A.q[1] : A.q[0]); a_q0_low = A.q[0] == 0; }
      ordered[1] = pads[np++];
    } else {       // bits (1,2)
      ordered[0] = slot_pos(A.q[0]);
      ordered[1] = A.nq == 2 ? slot_pos(A.q[1]) : pads[np++];
      a_q0_low = true;
    }
    if (fr.b >= 0) {
      const QGate &B = gates[fr.b];
      ordered[2] = slot_pos(B.q[0]);
      ordered[3] = B.nq == 2 ? slot_pos(B.q[1]) : pads[np++];
    } else {
      ordered[2] = pads[np++];
      ordered[3] = pads[np++];
    }
    build_round(R, ordered, w, kSB, true);
    R.sync = (r == seg_end[r] - 1);
    R.fast = (uint8_t)((a_has0 ? 3 : 1) + (fr.b >= 0 ? 1 : 0));
    R.gate[0] = (uint16_t)put_matrix(A, a_q0_low);
    R.gate[1] = fr.b >= 0 ? (uint16_t)put_matrix(gates[fr.b], true) : R.gate[0];
    R.ngates = (uint8_t)(1 + (fr.b >= 0));
  }
  if (s.selftest_host) {
    if (!s.plan_only) emulate_tile_pass<float>(p, s.selftest_host, nullptr, true);
    return leftover;
  }
  emit_tile_pass(s, p, TV_PIPE2_F32);
  return leftover;
}

// float states: partition into passes over 13-bit tiles (the 4 lowest global bits are always in the tile: 128-byte runs)
static int apply_gate_sequence_f32(State &s, const std::vector<QGate> &gates) {
  constexpr int kTBf = 13;
  PassPacker packer(gates);
  int passes = 0;
  while (!packer.finished()) {
    uint64_t Q = 0xF;
    const std::vector<int> sel = packer.select(Q, kTBf, kMaxTileGates, kMaxTileGates);
    for (int q = 0; q < s.nq && __builtin_popcountll(Q) < kTBf; q++) Q |= 1ull << q;
    std::vector<int> tile_bits;
    for (int q = 0; q < 64; q++)
      if ((Q >> q) & 1) tile_bits.push_back(q);
    const std::vector<int> back = run_tile_pass_f32(s, gates, sel, tile_bits);
    passes++;
    if (back.size() == sel.size()) throw Error("tile pass: no progress");
    packer.commit(sel, back);
  }
  return passes;
}

// ------------------------------------------------------------------------------------------ queue-level absorption
// A dense 1-qubit gate whose neighbour in time on that qubit is a dense (non-diagonal) 2-qubit gate is multiplied
// into it on the host: u(a) directly after G(a,b) becomes (u x 1) G, directly before it G (u x 1); consecutive 1-qubit
// gates on one qubit collapse the same way.  "Directly" = no other op (gate or per-state Pauli) on that qubit in
// between, so the reordering is exact.  Typical transpiled circuits (rz / sx / u around every cx) lose 2/3 of their ops
// and turn their generic rounds into fast ones.  This is the engine's counterpart of what Fusion::optimize_circuit
// does for 1-qubit gates when Aer's fusion is switched off in favour of the gate queue.  B200SV_QUEUE_ABSORB=0 disables.
static void mat4_mul(const cd_t *A, const cd_t *B, cd_t *C) {  // column-major 4x4, C = A B
  cd_t out[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      cd_t acc = 0;
      for (int k = 0; k < 4; k++) acc += A[r + 4 * k] * B[k + 4 * c];
      out[r + 4 * c] = acc;
    }
  std::copy(out, out + 16, C);
}
static void embed_1q(const cd_t *u /*column-major 2x2*/, bool on_bit0, cd_t *E) {
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      const int rl = r & 1, rh = r >> 1, cl = c & 1, ch = c >> 1;
      E[r + 4 * c] = on_bit0 ? (rh == ch ? u[rl + 2 * cl] : cd_t(0)) : (rl == cl ? u[rh + 2 * ch] : cd_t(0));
    }
}
static void absorb_one_qubit_gates(std::vector<QGate> &gates, int nq, std::vector<std::array<double, 32>> &store) {
  static const int env_on = [] { const char *e = getenv("B200SV_QUEUE_ABSORB"); return e ? atoi(e) : 1; }();
  if (!env_on) return;
  store.reserve(gates.size());  // pointers into `store` must stay valid
  auto own = [&](QGate &g) {    // give g a private, writable matrix
    cd_t *cur = reinterpret_cast<cd_t *>(const_cast<double *>(g.mat));
    for (auto &st : store)
      if (reinterpret_cast<cd_t *>(st.data()) == cur) return cur;
    store.emplace_back();
    std::copy(g.mat, g.mat + 32, store.back().data());
    g.mat = store.back().data();
    return reinterpret_cast<cd_t *>(store.back().data());
  };
  std::vector<int> last(nq, -1);  // index (in gates) of the last kept op on each qubit
  std::vector<char> dead(gates.size(), 0);
  auto two = [&](int i) { return i >= 0 && gates[i].mat && gates[i].nq == 2; };
  auto dense1 = [&](int i) { return i >= 0 && gates[i].mat && gates[i].nq == 1; };
  // a 1-qubit gate may join a 2-qubit one when that does not make a cheap diagonal gate dense
  auto joins = [&](int one, int twoq) { return two(twoq) && (!is_diag(gates[twoq]) || is_diag(gates[one])); };
  for (int i = 0; i < (int)gates.size(); i++) {
    QGate &g = gates[i];
    if (dense1(i)) {
      const int a = g.q[0], j = last[a];
      const cd_t *u = reinterpret_cast<const cd_t *>(g.mat);
      if (joins(i, j)) {          // (u on a) after G: G <- E G
        cd_t E[16];
        embed_1q(u, gates[j].q[0] == a, E);
        cd_t *G = own(gates[j]);
        mat4_mul(E, G, G);
        dead[i] = 1;
        continue;
      }
      if (dense1(j)) {            // u after v on the same qubit: v <- u v
        cd_t *v = own(gates[j]);
        cd_t out[4];
        for (int c = 0; c < 2; c++)
          for (int r = 0; r < 2; r++) out[r + 2 * c] = u[r] * v[2 * c] + u[r + 2] * v[1 + 2 * c];
        std::copy(out, out + 4, v);
        dead[i] = 1;
        continue;
      }
      last[a] = i;
      continue;
    }
    if (two(i)) {
      const int j0 = last[g.q[0]], j1 = last[g.q[1]];
      if (j0 == j1 && two(j0)) {
        // the previous op on BOTH qubits is a 2-qubit gate on the same pair: H <- G H (cx rz cx -> one gate)
        QGate &h = gates[j0];
        cd_t P[16];
        const cd_t *G = reinterpret_cast<const cd_t *>(g.mat);
        const bool same_order = h.q[0] == g.q[0];
        for (int c = 0; c < 4; c++)
          for (int r = 0; r < 4; r++) {
            const int rs = same_order ? r : ((r >> 1) | ((r & 1) << 1)), cs = same_order ? c : ((c >> 1) | ((c & 1) << 1));
            P[r + 4 * c] = G[rs + 4 * cs];
          }
        cd_t *H = own(h);
        mat4_mul(P, H, H);
        dead[i] = 1;
        continue;
      }
      for (int x = 0; x < 2; x++) {
        const int j = last[g.q[x]];
        if (dense1(j) && joins(j, i)) {  // u directly before G on this qubit: G <- G E
          cd_t E[16];
          embed_1q(reinterpret_cast<const cd_t *>(gates[j].mat), x == 0, E);
          cd_t *G = own(g);
          mat4_mul(G, E, G);
          dead[j] = 1;
        }
      }
    }
    for (int x = 0; x < g.nq; x++) last[g.q[x]] = i;
  }
  size_t w = 0;
  for (size_t i = 0; i < gates.size(); i++)
    if (!dead[i]) gates[w++] = gates[i];
  gates.resize(w);
}

// Partition an op sequence into tile passes (in-order greedy with dependency blocking) and run them.
// kind[i]: 1 = dense 1-qubit, 2 = dense 2-qubit, 3 = per-state Pauli on one qubit (code table slot[i]).
// Returns the number of HBM passes used.
int apply_gate_sequence(State &s, int ngates, const int *nq, const uint64_t *qubits, const double *mats,
                        int low_bits, const int *slot, const uint8_t *codes_host, int nslots) {
  std::vector<QGate> gates(ngates);
  bool any_pauli = false;
  for (int i = 0; i < ngates; i++) {
    if (nq[i] < 1 || nq[i] > 3) throw Error("apply_gate_sequence: op kind must be 1, 2 (dense) or 3 (per-state Pauli)");
    const bool pauli = nq[i] == 3;
    gates[i].nq = pauli ? 1 : nq[i];
    for (int j = 0; j < gates[i].nq; j++) {
      if (qubits[2 * i + j] >= (uint64_t)s.nq) throw Error("apply_gate_sequence: qubit out of range");
      gates[i].q[j] = (int)qubits[2 * i + j];
    }
    if (nq[i] == 2 && gates[i].q[0] == gates[i].q[1]) throw Error("apply_gate_sequence: duplicate qubit");
    gates[i].mat = pauli ? nullptr : mats + 32 * (size_t)i;
    gates[i].slot = 0;
    if (pauli) {
      if (!slot || !codes_host || slot[i] < 0 || slot[i] >= nslots) throw Error("apply_gate_sequence: bad Pauli slot");
      gates[i].slot = slot[i];
      any_pauli = true;
    }
  }
  // Circuits rich in commuting diagonal gates (QFT's controlled phases, cz / rz layers): the engine's fusion pass
  // (planner.cu, commutation aware) regroups the queue into small dense blocks and WIDE diagonal blocks; a wide
  // diagonal block is one streaming pass (gates.cu: diag_layer_kernel) instead of one tile-round slot per gate, the
  // dense blocks in between go through the tile passes as before.  The role Fusion::optimize_circuit's diagonal
  // fusion plays in the reference (src/transpile/fusion.hpp:816-821), without its 2^k table limit.
  static const int env_layers = [] { const char *e = getenv("B200SV_QUEUE_DIAG_LAYERS"); return e ? atoi(e) : 1; }();
  if (env_layers && !s.in_layer_split && !any_pauli && !s.selftest_host && !s.capture && s.nstates == 1 && s.nq >= 16 &&
      s.global_nq <= s.nq) {
    int ndiag2 = 0;
    std::vector<uint8_t> isd(ngates);
    bool layer_ok = true;
    for (int i = 0; i < ngates; i++) {
      isd[i] = is_diag(gates[i]) ? 1 : 0;
      if (isd[i] && gates[i].nq == 2) ndiag2++;
      if (isd[i]) {
        const int dim = 1 << gates[i].nq;
        for (int d = 0; d < dim; d++)
          if (gates[i].mat[2 * (d + dim * d)] == 0.0 && gates[i].mat[2 * (d + dim * d) + 1] == 0.0) layer_ok = false;  // projector
      }
    }
    if (layer_ok && ndiag2 >= 24 && 3 * ndiag2 >= ngates) {
      std::vector<int> off(ngates + 1, 0), qs, blk(ngates);
      for (int i = 0; i < ngates; i++) {
        for (int j = 0; j < gates[i].nq; j++) qs.push_back(gates[i].q[j]);
        off[i + 1] = (int)qs.size();
      }
      int nblocks = 0;
      static const int env_dense = [] { const char *e = getenv("B200SV_LAYER_DENSE_QUBITS"); return e ? std::min(12, std::max(2, atoi(e))) : 5; }();
      fuse_assign(ngates, off.data(), qs.data(), isd.data(), env_dense, 64, 62, blk.data(), &nblocks);
      std::vector<std::vector<int>> members(nblocks);
      for (int i = 0; i < ngates; i++) members[blk[i]].push_back(i);
      int passes = 0;
      std::vector<int> seq_nq;
      std::vector<uint64_t> seq_q;
      std::vector<double> seq_m;
      auto flush_seq = [&] {
        if (seq_nq.empty()) return;
        s.in_layer_split = true;
        try {
          passes += apply_gate_sequence(s, (int)seq_nq.size(), seq_nq.data(), seq_q.data(), seq_m.data(), low_bits);
        } catch (...) { s.in_layer_split = false; throw; }
        s.in_layer_split = false;
        seq_nq.clear(); seq_q.clear(); seq_m.clear();
      };
      for (int b = 0; b < nblocks; b++) {
        bool all_diag = true;
        for (int i : members[b]) all_diag = all_diag && isd[i];
        if (all_diag && members[b].size() >= 6) {
          flush_seq();
          std::vector<int> lnq;
          std::vector<uint64_t> lq;
          std::vector<double> ld;
          for (int i : members[b]) {
            const QGate &g = gates[i];
            const int dim = 1 << g.nq;
            lnq.push_back(g.nq);
            lq.push_back((uint64_t)g.q[0]);
            lq.push_back(g.nq == 2 ? (uint64_t)g.q[1] : 0);
            for (int d = 0; d < 4; d++) {
              ld.push_back(d < dim ? g.mat[2 * (d + dim * d)] : 0.0);
              ld.push_back(d < dim ? g.mat[2 * (d + dim * d) + 1] : 0.0);
            }
          }
          launch_diag_layer(s, (int)lnq.size(), lnq.data(), lq.data(), ld.data());
          passes++;
          continue;
        }
        for (int i : members[b]) {
          seq_nq.push_back(gates[i].nq);
          seq_q.push_back((uint64_t)gates[i].q[0]);
          seq_q.push_back(gates[i].nq == 2 ? (uint64_t)gates[i].q[1] : 0);
          seq_m.insert(seq_m.end(), gates[i].mat, gates[i].mat + 32);
        }
      }
      flush_seq();
      return passes;
    }
  }
  std::vector<std::array<double, 32>> merged;
  if ((s.precision == B200SV_F64 && s.nq >= 12) || (s.precision == B200SV_F32 && s.nq >= 13)) {
    absorb_one_qubit_gates(gates, s.nq, merged);
    ngates = (int)gates.size();
  }
  // B200SV_TILE_BITS = 11 | 12 selects the tile size (default 12)
  static const int env_tb = [] { const char *e = getenv("B200SV_TILE_BITS"); return e ? atoi(e) : 0; }();
  static const int env_pipe3 = [] { const char *e = getenv("B200SV_TILE_PIPE"); return e ? atoi(e) == 3 : false; }();
  const int kTB = ((env_tb == 11 || env_pipe3) && !s.selftest_host) ? 11 : 12;
  static const int env_f32 = [] { const char *e = getenv("B200SV_TILE_F32"); return e ? atoi(e) : 1; }();
  if (s.precision == B200SV_F32 && s.nq >= 13 && !any_pauli && env_f32) return apply_gate_sequence_f32(s, gates);
  const bool tiled = s.precision == B200SV_F64 && s.nq >= kTB;
  if (!tiled && s.selftest_host) throw Error("selftest: this state size / precision / op mix does not take the tile passes");
  if (!tiled) {  // small or single-precision states: one streaming pass per op
    for (auto &g : gates) {
      if (g.mat && s.capture) {
        s.capture->passes.emplace_back();
        TilePlannedPass &pp = s.capture->passes.back();
        pp.variant = TV_DENSE;
        pp.dnq = g.nq;
        pp.dq[0] = g.q[0]; pp.dq[1] = g.nq == 2 ? g.q[1] : 0;
        std::copy(g.mat, g.mat + 32, pp.dmat);
        continue;
      }
      if (g.mat) { launch_dense(s, g.q, g.nq, nullptr, 0, g.mat); continue; }
      if (s.capture) throw Error("tile plan: per-state Pauli ops cannot be captured");
      std::vector<uint64_t> m4(4 * (size_t)s.nstates, 0);  // batched_pauli_func masks (qubitvector_thrust.hpp:2819)
      for (int64_t st = 0; st < s.nstates; st++) {
        const int code = codes_host[(size_t)g.slot * s.nstates + st];
        const uint64_t bit = 1ull << g.q[0];
        m4[4 * st] = (code == 1 || code == 2) ? bit : 0;
        m4[4 * st + 1] = (code == 2 || code == 3) ? bit : 0;
        m4[4 * st + 2] = code == 2;
        m4[4 * st + 3] = code != 0;
      }
      launch_batched_pauli(s, m4.data());
    }
    return ngates;
  }
  uint8_t *dev_codes = nullptr;
  if (any_pauli && s.selftest_host) s.selftest_codes = codes_host;
  if (any_pauli && !s.selftest_host) {
    const size_t bytes = (size_t)nslots * s.nstates;
    void *hm = s.ensure_pinned(bytes);
    // + slack: the pipelined kernel's idle group reads (and discards) codes of up to 2 * grid tiles past the end
    dev_codes = (uint8_t *)s.ensure_scratch(bytes + 1024);
    B200_CUDA(cudaStreamSynchronize(s.stream));
    memcpy(hm, codes_host, bytes);
    B200_CUDA(cudaMemcpyAsync(dev_codes, hm, bytes, cudaMemcpyHostToDevice, s.stream));
  }
  low_bits = std::max(1, std::min(low_bits, 5));
  // tuning knobs (read once): B200SV_TILE_MAX_GATES caps the dense gates riding on one pass (FP64/HBM balance),
  // B200SV_TILE_LOW_BITS the number of low global bits forced into every tile (coalescing run length)
  static const int env_max_gates = [] { const char *e = getenv("B200SV_TILE_MAX_GATES"); return e ? atoi(e) : 0; }();
  static const int env_low_bits = [] { const char *e = getenv("B200SV_TILE_LOW_BITS"); return e ? atoi(e) : 0; }();
  const int max_gates = env_max_gates > 0 ? std::min(env_max_gates, kMaxTileGates) : kMaxTileGates;
  if (env_low_bits > 0) low_bits = std::min(env_low_bits, 5);
  static const int env_pack = [] { const char *e = getenv("B200SV_TILE_PACK"); return e ? atoi(e) : 1; }();
  std::vector<int> rem(ngates);
  for (int i = 0; i < ngates; i++) rem[i] = i;
  PassPacker packer(gates);
  int passes = 0;
  while (env_pack ? !packer.finished() : !rem.empty()) {
    uint64_t Q = (1ull << low_bits) - 1, blocked = 0;
    std::vector<int> sel, rest;
    if (env_pack) {
      sel = packer.select(Q, kTB, max_gates, 4 * kMaxTileGates);
    } else {
      int ndense = 0;
      for (int i : rem) {
        const uint64_t m = qmask(gates[i]);
        const bool dense = gates[i].mat != nullptr;
        if ((m & blocked) || (dense && ndense >= max_gates) || (int)sel.size() >= 4 * kMaxTileGates) {
          blocked |= m; rest.push_back(i); continue;
        }
        if (__builtin_popcountll(Q | m) <= kTB) { Q |= m; sel.push_back(i); ndense += dense; }
        else { blocked |= m; rest.push_back(i); }
      }
    }
    // fill the tile with the lowest unused global bits
    for (int q = 0; q < s.nq && __builtin_popcountll(Q) < kTB; q++) Q |= 1ull << q;
    std::vector<int> tile_bits;
    for (int q = 0; q < 64; q++)
      if ((Q >> q) & 1) tile_bits.push_back(q);
    const std::vector<int> back = run_tile_pass(s, gates, sel, tile_bits, dev_codes, kTB);
    passes++;
    static const int env_trace = [] { const char *e = getenv("B200SV_TILE_TRACE"); return e ? atoi(e) : 0; }();
    if (env_trace) fprintf(stderr, "b200sv tile pass %d: %zu ops (%zu deferred)\n", passes, sel.size(), back.size());
    if (env_pack) {
      if (back.size() == sel.size()) throw Error("tile pass: no progress");
      packer.commit(sel, back);
      continue;
    }
    if (!back.empty()) {  // deferred ops commute with everything earlier that is still queued: program order is safe
      rest.insert(rest.end(), back.begin(), back.end());
      std::sort(rest.begin(), rest.end());
    }
    rem.swap(rest);
  }
  return passes;
}

// ------------------------------------------------------------------------------------------ plan API (sharded executor)
TilePlan *tile_plan_build(State &s, int ngates, const int *nq, const uint64_t *qubits, const double *mats) {
  TilePlan *plan = new TilePlan();
  s.capture = plan;
  try {
    if (ngates > 0) apply_gate_sequence(s, ngates, nq, qubits, mats, 3);
  } catch (...) {
    s.capture = nullptr;
    delete plan;
    throw;
  }
  s.capture = nullptr;
  return plan;
}
void tile_plan_free(TilePlan *plan) { delete plan; }
int tile_plan_passes(const TilePlan *plan) { return (int)plan->passes.size(); }
// global index bits a pass addresses inside its tiles (a slab may only fix bits outside this mask); ~0 for passes
// that cannot be restricted to a slab
uint64_t tile_plan_pass_mask(const TilePlan *plan, int i) {
  const TilePlannedPass &pp = plan->passes.at(i);
  if (pp.variant == TV_DENSE) return ~0ull;
  uint64_t m = 0;
  const int shift = pp.variant == TV_PIPE2_F32 ? 1 : 0;  // float passes address 16-byte slots: position - 1, qubit 0 inside
  for (int u = 0; u < pp.p.ins.n; u++) m |= 1ull << (pp.p.ins.pos[u] + shift);
  if (shift) m |= 1ull;
  return m;
}
// host evaluation of a captured small-state pass (scheduler / executor self-tests only)
template <typename T>
static void emulate_dense_gate(void *host, int nq, const TilePlannedPass &pp) {
  typedef std::complex<T> C;
  C *psi = reinterpret_cast<C *>(host);
  const cd_t *M = reinterpret_cast<const cd_t *>(pp.dmat);  // column major
  const int k = pp.dnq, dim = 1 << k;
  for (uint64_t i = 0; i < (1ull << nq); i++) {
    bool lead = true;
    for (int b = 0; b < k; b++) lead = lead && !((i >> pp.dq[b]) & 1);
    if (!lead) continue;
    cd_t x[4], y[4];
    uint64_t idx[4];
    for (int e = 0; e < dim; e++) {
      idx[e] = i;
      for (int b = 0; b < k; b++)
        if ((e >> b) & 1) idx[e] |= 1ull << pp.dq[b];
      x[e] = cd_t(psi[idx[e]]);
    }
    for (int r = 0; r < dim; r++) {
      y[r] = 0;
      for (int c = 0; c < dim; c++) y[r] += M[r + dim * c] * x[c];
    }
    for (int e = 0; e < dim; e++) psi[idx[e]] = C(y[e]);
  }
}

void tile_plan_launch(State &s, const TilePlan *plan, int i, const SlabSpec *slab) {
  const TilePlannedPass &pp = plan->passes.at(i);
  const bool has_slab = slab && slab->nbits > 0;
  if (pp.variant == TV_DENSE) {
    if (has_slab) throw Error("tile plan: this pass cannot be restricted to a slab");
    if (s.selftest_host) {
      if (s.precision == B200SV_F64) emulate_dense_gate<double>(s.selftest_host, s.nq, pp);
      else emulate_dense_gate<float>(s.selftest_host, s.nq, pp);
      return;
    }
    launch_dense(s, pp.dq, pp.dnq, nullptr, 0, pp.dmat);
    return;
  }
  const TilePassParams *use = &pp.p;
  uint64_t offset = 0;  // in amplitudes
  static thread_local TilePassParams q;
  if (has_slab) {
    q = pp.p;
    const int shift = pp.variant == TV_PIPE2_F32 ? 1 : 0;
    std::vector<int> pos(q.ins.pos, q.ins.pos + q.ins.n);
    for (int b = 0; b < slab->nbits; b++) {
      const int sp = slab->pos[b];
      if (sp < shift || sp >= s.nq) throw Error("tile plan: slab bit out of range");
      if (std::find(pos.begin(), pos.end(), sp - shift) != pos.end()) throw Error("tile plan: slab bit lies inside the pass's tile");
      pos.push_back(sp - shift);
      if ((slab->value >> b) & 1u) offset |= 1ull << sp;
    }
    std::sort(pos.begin(), pos.end());
    if ((int)pos.size() > kMaxInsert) throw Error("tile plan: too many fixed bits");
    q.ins.n = (int)pos.size();
    for (size_t u = 0; u < pos.size(); u++) q.ins.pos[u] = (uint8_t)pos[u];
    q.ntiles = pp.p.ntiles >> slab->nbits;
    use = &q;
  }
  if (s.selftest_host) {  // executor self-test: interpret the (slab-patched) parameter block on the host slice
    char *hb = (char *)s.selftest_host + offset * s.amp_bytes();
    if (pp.variant == TV_PIPE2_F32) emulate_tile_pass<float>(*use, hb, nullptr, true);
    else if (pp.variant == TV_PASS11 || pp.variant == TV_PIPE3_FAST) throw Error("tile plan: 2^11 tiles are not emulated");
    else emulate_tile_pass<double>(*use, hb, nullptr, false);
    return;
  }
  char *base = (char *)s.data + offset * s.amp_bytes();
  launch_tile_variant(s, *use, pp.variant, (double2 *)base, slab ? slab->sm_limit : 0);
}

#ifdef B200SV_TILE_PROFILE
extern "C" void b200sv_tile_profile(unsigned long long *out4, int reset) {
  cudaMemcpyFromSymbol(out4, g_tile_prof, sizeof(g_tile_prof));
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(g_tile_prof, z, sizeof(z)); }
}
#endif

}  // namespace b200sv
