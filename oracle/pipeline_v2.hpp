// TEST INFRASTRUCTURE ONLY — CPU restatement of the V2 pipeline stages around the saturation analyzers:
//   CostAwareOptimizer.Optimize   internal/engines/pipeline/cost_aware_optimizer.go:39-197
//   Enforcer.EnforcePolicy        internal/engines/pipeline/enforcer.go:55-183
//   SaturationAnalyzer.Analyze    internal/engines/analyzers/saturation_v2/analyzer.go:59-138 (the arithmetic; the
//                                 rolling k2 history and the capacity store stay with the caller)
// Name-keyed maps become index-keyed arrays in the caller's slice order; `sort.Slice` (unstable) is a stable sort
// here — one of the orders the reference can produce.  Never linked into the product.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace oracle_v2 {

// float64 -> int as Go does on amd64 (CVTTSD2SQ): out-of-range and NaN give the "integer indefinite" value
static inline long long go_int(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return (long long)0x8000000000000000ull;
  return (long long)x;
}

// One model.  current/cost/cap: per variant (slice order).  targets out.  cost_aware_optimizer.go:54-61
static inline void cost_aware_model(int V, double required, double spare, const int* current, const double* cost,
                                    const double* cap, int* target) {
  for (int v = 0; v < V; v++) target[v] = current[v];                      // initTargets :182-188
  std::vector<int> order(V);
  for (int v = 0; v < V; v++) order[v] = v;
  if (required > 0) {                                                      // costAwareScaleUp :75-104
    auto eff = [&](int v) { return cap[v] <= 0 ? 1.79769313486231570814527423731704357e+308 : cost[v] / cap[v]; };   // :233-238
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return eff(a) < eff(b); });
    double remaining = required;
    for (int k = 0; k < V; k++) {
      const int v = order[k];
      if (remaining <= 0) break;
      if (cap[v] <= 0) continue;
      const long long need = go_int(std::ceil(remaining / cap[v]));
      target[v] = (int)((long long)target[v] + need);
      remaining -= (double)need * cap[v];
    }
  } else if (spare > 0) {                                                  // costAwareScaleDown :111-168
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    int cheapest = -1;                                                     // findCheapestVariant :191-201
    double min_cost = 1.79769313486231570814527423731704357e+308;
    for (int v = 0; v < V; v++) if (cost[v] < min_cost) { min_cost = cost[v]; cheapest = v; }
    double remaining = spare;
    for (int k = 0; k < V; k++) {
      const int v = order[k];
      if (remaining <= 0) break;
      if (cap[v] <= 0) continue;
      const int cur = target[v];
      int min_rep = 0;
      if (v == cheapest) {
        bool other = false;
        for (int u = 0; u < V; u++) if (u != cheapest && target[u] > 0) { other = true; break; }
        if (!other) min_rep = 1;
      }
      const int removable = cur - min_rep;
      if (removable <= 0) continue;
      long long rem = go_int(std::floor(remaining / cap[v]));
      if (rem > removable) rem = removable;
      if (rem <= 0) continue;
      target[v] = cur - (int)rem;
      remaining -= (double)rem * cap[v];
    }
  }
}

// One model.  target in/out (-1 = the variant is not in the targets map).  Returns `applied`.  enforcer.go:55-183
static inline bool enforce_model(int V, int* target, const double* cost, const unsigned char* has_cost, bool s2z_enabled,
                                 double request_count, bool request_error, const int* name_rank = nullptr) {
  if (s2z_enabled) {                                                       // applyScaleToZero :86-127
    if (request_error || request_count > 0) return false;
    for (int v = 0; v < V; v++) if (target[v] >= 0) target[v] = 0;
    return true;
  }
  long long total = 0;                                                     // ensureMinimumReplicas :130-183
  for (int v = 0; v < V; v++) if (target[v] >= 0) total += target[v];
  if (total > 0) return false;
  int cheapest = -1, cheapest_rank = -1;
  double cheapest_cost = -1.0;
  for (int v = 0; v < V; v++) {
    if (target[v] < 0) continue;
    const double c = (has_cost && !has_cost[v]) ? 10.0 : cost[v];          // saturation.DefaultVariantCost
    const int rk = name_rank ? name_rank[v] : v;                           // `variant < cheapestVariant` compares names (:161)
    if (cheapest_cost < 0 || c < cheapest_cost || (c == cheapest_cost && rk < cheapest_rank)) { cheapest = v; cheapest_rank = rk; cheapest_cost = c; }
  }
  if (cheapest >= 0) { target[cheapest] = 1; return true; }
  return false;
}


// ---- SaturationAnalyzer.Analyze (saturation_v2/analyzer.go:59-138), one model ------------------------------------------
struct V2ModelIn {
  int V;                                  // variants, VariantStates order
  const int* vro;                         // [V+1] replica ranges (absolute indices into the rep_* arrays)
  const long long *rep_total_kv, *rep_tokens_in_use, *rep_queue_len, *rep_k2;
  const double *rep_avg_in, *rep_avg_out, *rep_hit;
  const int* slice_order;                 // replica indices of the model in ReplicaMetrics order (or nullptr: grouped order)
  const int *var_current, *var_pending;
  const double* var_fallback;
  double kv_threshold, scale_up_threshold, scale_down_boundary;
  bool has_queue; long long queue_size, queue_bytes;
};
struct V2ModelOut {
  long long *rep_k1, *rep_effective, *rep_demand; unsigned char* rep_saturated;        // absolute replica indices
  int* var_ready; double *var_cap, *var_total_cap, *var_total_demand, *var_util;       // [V]
  double total_supply, total_demand, utilization, required, spare;
};

static inline long long median_i64(std::vector<long long> v) {             // analyzer.go:505-519
  const size_t n = v.size();
  if (n == 0) return 0;
  std::stable_sort(v.begin(), v.end());
  if (n % 2 == 0) return (v[n / 2 - 1] + v[n / 2]) / 2;
  return v[n / 2];
}

static inline void saturation_v2_model(const V2ModelIn& in, V2ModelOut& out) {
  double total_supply = 0, total_anticipated = 0, total_demand = 0;
  const int r_begin = in.vro[0], r_end = in.vro[in.V];
  for (int v = 0; v < in.V; v++) {
    std::vector<long long> caps;
    double demand_sum = 0;
    for (int r = in.vro[v]; r < in.vro[v + 1]; r++) {
      long long k1 = 0, eff = 0, demand = 0; bool sat = false;
      if (in.rep_total_kv[r] > 0) {                                          // computeReplicaCapacity :142-211
        demand = in.rep_tokens_in_use[r];
        if (in.rep_avg_in[r] > 0) demand += in.rep_queue_len[r] * go_int(in.rep_avg_in[r]);
        k1 = go_int((double)in.rep_total_kv[r] * in.kv_threshold);
        const long long k2 = in.rep_k2[r] < 0 ? k1 : in.rep_k2[r];           // computeK2 priority 4: fall back to k1
        eff = k2 < k1 ? k2 : k1;
        sat = demand >= eff;
        caps.push_back(eff);
        demand_sum += (double)demand;
      }
      if (out.rep_k1) out.rep_k1[r] = k1;
      if (out.rep_effective) out.rep_effective[r] = eff;
      if (out.rep_demand) out.rep_demand[r] = demand;
      if (out.rep_saturated) out.rep_saturated[r] = sat ? 1 : 0;
    }
    int ready = in.var_current[v] - in.var_pending[v];                       // aggregateByVariant :300-303
    if (ready < 0) ready = 0;
    double cap = 0;
    if (!caps.empty()) cap = (double)median_i64(caps);
    else cap = in.var_fallback[v];
    const double total_cap = (double)ready * cap;
    double util = 0;
    if (total_cap > 0) util = demand_sum / total_cap;
    out.var_ready[v] = ready; out.var_cap[v] = cap; out.var_total_cap[v] = total_cap;
    out.var_total_demand[v] = demand_sum; out.var_util[v] = util;
    total_supply += total_cap;                                               // Analyze :88-96
    total_demand += demand_sum;
    total_anticipated += (double)(ready + in.var_pending[v]) * cap;
  }
  if (in.has_queue && !(in.queue_size == 0 && in.queue_bytes == 0)) {        // estimateSchedulerQueueDemand :471-501
    double ai = 0, ao = 0, ah = 0; int cnt = 0;                              // computeModelWorkloadAverages :438-455
    for (int k = 0; k < r_end - r_begin; k++) {
      const int r = in.slice_order ? in.slice_order[r_begin + k] : r_begin + k;
      if (in.rep_avg_in[r] > 0 || in.rep_avg_out[r] > 0) { ai += in.rep_avg_in[r]; ao += in.rep_avg_out[r]; ah += in.rep_hit[r]; cnt++; }
    }
    if (cnt > 0) { ai /= (double)cnt; ao /= (double)cnt; ah /= (double)cnt; }
    const double from_bytes = (double)in.queue_bytes / 4.0;                  // BytesPerToken
    const double from_count = (double)in.queue_size * ai;
    double input_tokens = from_bytes;
    if (from_count > input_tokens) input_tokens = from_count;
    input_tokens *= (1 - ah);
    const double output_tokens = (double)in.queue_size * ao;
    total_demand += input_tokens + output_tokens;
  }
  double utilization = 0;
  if (total_supply > 0) utilization = total_demand / total_supply;
  double required = 0, spare = 0;
  if (in.scale_up_threshold > 0) required = total_demand / in.scale_up_threshold - total_anticipated;
  if (required < 0) required = 0;
  if (in.scale_down_boundary > 0) spare = total_supply - total_demand / in.scale_down_boundary;
  if (spare < 0) spare = 0;
  out.total_supply = total_supply; out.total_demand = total_demand; out.utilization = utilization;
  out.required = required; out.spare = spare;
}

// estimateCapacityFromParams (analyzer.go:418-437): the k2 derivation the caller's priority chain uses
static inline long long estimate_capacity_from_params(long long max_batched_tokens, long long max_num_seqs, double avg_in,
                                                      double avg_out) {
  if (max_batched_tokens <= 0 || avg_out <= 0) return 0;
  const double B = (double)max_batched_tokens, S = (double)max_num_seqs;
  double n_steady = B * avg_out / (avg_in + avg_out);
  if (n_steady > S) n_steady = S;
  const long long k2 = go_int(n_steady * (avg_in + avg_out / 2));
  return k2 > 0 ? k2 : 0;
}

}  // namespace oracle_v2
