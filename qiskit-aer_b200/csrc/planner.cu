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
