// b200sv epoch planner for sharded registers (host code only).
//
// Role in the reference: CacheBlocking (src/transpile/cacheblocking.hpp, which reorders the circuit and inserts
// swap_chunk ops so that every gate acts on "blocked" = chunk-local qubits) together with
// ParallelStateExecutor::apply_chunk_swap (src/simulators/parallel_state_executor.hpp:1134).  Design here (DESIGN.md
// section 5): one slice per GPU, a logical -> physical qubit map instead of swap-backs, and EPOCH scheduling -- run every
// op that is executable under the current map (respecting dependencies through shared qubits), then bring in the
// global qubits the blocked ops wait for, evicting the local qubits whose next use is farthest away (Belady), repeat.
// Diagonal ops and controls never block on a global qubit (they are resolved from the chunk index).
#include "common.cuh"
#include <complex>

namespace b200sv {

// plan encoding (int64): op   -> 0, op index, nq, physical qubits...
//                        swap -> 1, local position, global bit
//                        all-to-all swap of k pairs -> 2, k, local positions..., global bits...
void plan_epochs(int n, int nl, int gbits, int nops, const int *op_off, const int *op_qubits,
                 const uint8_t *need_local, int min_run_bits, bool multi_swap, int *phys, std::vector<int64_t> &out) {
  std::vector<int> remaining(nops);
  for (int i = 0; i < nops; i++) remaining[i] = i;
  while (!remaining.empty()) {
    std::vector<char> blocked(n, 0), frontier(n, 0), is_wanted(n, 0);
    std::vector<int> rest, wanted;
    for (int i : remaining) {
      const int b = op_off[i], e = op_off[i + 1];
      bool hit = false, glob = false;
      for (int k = b; k < e; k++) hit = hit || blocked[op_qubits[k]];
      if (!hit)
        for (int k = b; k < e; k++) glob = glob || (need_local[k] && phys[op_qubits[k]] >= nl);
      if (hit || glob) {
        for (int k = b; k < e; k++) blocked[op_qubits[k]] = 1;
        rest.push_back(i);
        if (!hit) {
          for (int k = b; k < e; k++) frontier[op_qubits[k]] = 1;  // never evict a partner of the op we swap for
          for (int k = b; k < e; k++) {
            const int q = op_qubits[k];
            if (need_local[k] && phys[q] >= nl && !is_wanted[q]) { is_wanted[q] = 1; wanted.push_back(q); }
          }
        }
        continue;
      }
      out.push_back(0);
      out.push_back(i);
      out.push_back(e - b);
      for (int k = b; k < e; k++) out.push_back(phys[op_qubits[k]]);
    }
    remaining.swap(rest);
    if (remaining.empty()) break;
    // first use (position in `remaining`) of every logical qubit that must be local there
    std::vector<int64_t> first_use(n, (int64_t)1 << 60);
    for (int pos = (int)remaining.size() - 1; pos >= 0; pos--) {
      const int i = remaining[pos];
      for (int k = op_off[i]; k < op_off[i + 1]; k++)
        if (need_local[k]) first_use[op_qubits[k]] = pos;
    }
    std::vector<char> busy(n, 0);
    for (int q = 0; q < n; q++) busy[q] = is_wanted[q] || frontier[q];
    std::vector<int> inv(n);
    std::vector<std::pair<int, int>> pairs;
    const int nswap = std::min<int>((int)wanted.size(), gbits);
    for (int w = 0; w < nswap; w++) {
      const int q = wanted[w];
      for (int x = 0; x < n; x++) inv[phys[x]] = x;
      bool any_high = false;
      for (int p = nl - 1; p >= 0; p--) any_high = any_high || (!busy[inv[p]] && p >= min_run_bits);
      int best = -1;
      int64_t best_next = -1;
      for (int p = nl - 1; p >= 0; p--) {
        if (busy[inv[p]] || (any_high && p < min_run_bits)) continue;
        if (first_use[inv[p]] > best_next) { best = p; best_next = first_use[inv[p]]; }
      }
      if (best < 0) throw Error("plan_epochs: no local qubit available to evict");
      const int victim = inv[best];
      busy[victim] = 1;
      const int lpos = phys[victim], gpos = phys[q];
      pairs.push_back({lpos, gpos - nl});
      phys[victim] = gpos;
      phys[q] = lpos;
    }
    if (pairs.size() > 1 && multi_swap) {
      out.push_back(2);
      out.push_back((int64_t)pairs.size());
      for (auto &pr : pairs) out.push_back(pr.first);
      for (auto &pr : pairs) out.push_back(pr.second);
    } else {
      for (auto &pr : pairs) { out.push_back(1); out.push_back(pr.first); out.push_back(pr.second); }
    }
  }
}

}  // namespace b200sv

// ------------------------------------------------------------------------------------------ gate fusion (host)
// The engine's own fusion pass (role of Fusion::optimize_circuit, src/transpile/fusion.hpp:849, cost model
// :1002-1136) with the B200 cost model of DESIGN.md section 3: dense blocks grow to max_qubit (4 in double
// precision: the last size that is still HBM bound), purely diagonal blocks to max_diag_qubit (their table streams
// from shared memory / L2), and diagonal gates commute with each other, so a diagonal gate only depends on the
// last NON-diagonal block that shares a qubit with it.  List scheduling over open blocks: a gate may join block B
// iff no block emitted after B conflicts with it; preference: the last block it depends on, then any later block that
// stays within the size limit, else a new block.
namespace b200sv {

void fuse_assign(int nops, const int *op_off, const int *op_qubits, const uint8_t *op_is_diag, int max_qubit,
                 int window, int max_diag_qubit, int *block_of_op, int *nblocks_out) {
  struct Blk { uint64_t mask = 0; bool diag = true; };
  std::vector<Blk> blocks;
  for (int i = 0; i < nops; i++) {
    uint64_t qs = 0;
    for (int k = op_off[i]; k < op_off[i + 1]; k++) qs |= 1ull << op_qubits[k];
    const bool gdiag = op_is_diag[i] != 0;
    const int nb = (int)blocks.size(), lo = std::max(0, nb - window);
    int last_dep = -1;
    for (int b = nb - 1; b >= lo; b--)
      if ((qs & blocks[b].mask) && !(gdiag && blocks[b].diag)) { last_dep = b; break; }
    if (lo > 0 && last_dep < 0) last_dep = lo - 1;  // cannot prove independence from blocks outside the window
    int target = -1;
    for (int pass = 0; pass < 2 && target < 0; pass++) {
      const int b0 = pass == 0 ? last_dep : std::max(last_dep + 1, lo), b1 = pass == 0 ? last_dep + 1 : nb;
      if (pass == 0 && last_dep < lo) continue;
      for (int b = b0; b < b1; b++) {
        const int uni = __builtin_popcountll(blocks[b].mask | qs);
        const bool ok = (blocks[b].diag && gdiag) ? uni <= std::max(max_diag_qubit, max_qubit) : uni <= max_qubit;
        if (ok) { target = b; break; }
      }
    }
    if (target < 0) { blocks.emplace_back(); target = (int)blocks.size() - 1; }
    blocks[target].mask |= qs;
    blocks[target].diag = blocks[target].diag && gdiag;
    block_of_op[i] = target;
  }
  *nblocks_out = (int)blocks.size();
}

// product of the gates of one block on `k` block qubits (bit i of the matrix index <-> block_qubits[i]); gate g acts
// on gate_qubits[gate_off[g]..], its matrix is row-major 2^m x 2^m at gate_mats + gate_moff[g] (complex pairs).
// diag != 0: out = 2^k diagonal entries; else out = 2^k x 2^k row-major.
void fuse_block_matrix(int k, const int *block_qubits, int ngates, const int *gate_off, const int *gate_qubits,
                       const int64_t *gate_moff, const double *gate_mats, int diag, double *out) {
  typedef std::complex<double> cd;
  const size_t dim = (size_t)1 << k;
  cd *O = reinterpret_cast<cd *>(out);
  auto pos_of = [&](int q) {
    for (int i = 0; i < k; i++)
      if (block_qubits[i] == q) return i;
    throw Error("fuse_block_matrix: gate qubit outside the block");
  };
  if (diag) {
    for (size_t i = 0; i < dim; i++) O[i] = 1.0;
    for (int g = 0; g < ngates; g++) {
      const int m = gate_off[g + 1] - gate_off[g];
      const cd *U = reinterpret_cast<const cd *>(gate_mats) + gate_moff[g];
      int pos[16];
      for (int j = 0; j < m; j++) pos[j] = pos_of(gate_qubits[gate_off[g] + j]);
      const size_t gd = (size_t)1 << m;
      for (size_t i = 0; i < dim; i++) {
        size_t sub = 0;
        for (int j = 0; j < m; j++) sub |= ((i >> pos[j]) & 1) << j;
        O[i] *= U[sub * gd + sub];
      }
    }
    return;
  }
  std::vector<cd> M(dim * dim, 0.0), T(dim * dim);
  for (size_t i = 0; i < dim; i++) M[i * dim + i] = 1.0;
  for (int g = 0; g < ngates; g++) {  // M <- embed(U) M : rows of M mix inside each 2^m group
    const int m = gate_off[g + 1] - gate_off[g];
    const cd *U = reinterpret_cast<const cd *>(gate_mats) + gate_moff[g];
    int pos[16];
    for (int j = 0; j < m; j++) pos[j] = pos_of(gate_qubits[gate_off[g] + j]);
    const size_t gd = (size_t)1 << m;
    for (size_t r = 0; r < dim; r++) {
      size_t sub_r = 0, base = r;
      for (int j = 0; j < m; j++) { sub_r |= ((r >> pos[j]) & 1) << j; base &= ~((size_t)1 << pos[j]); }
      for (size_t c = 0; c < dim; c++) {
        cd acc = 0;
        for (size_t e = 0; e < gd; e++) {
          size_t src = base;
          for (int j = 0; j < m; j++) src |= ((e >> j) & 1) << pos[j];
          acc += U[sub_r * gd + e] * M[src * dim + c];
        }
        T[r * dim + c] = acc;
      }
    }
    M.swap(T);
  }
  std::copy(M.begin(), M.end(), O);
}

}  // namespace b200sv
