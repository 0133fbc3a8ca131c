// ORACLE — TEST INFRASTRUCTURE ONLY (see analyzer.hpp header).
// CPU restatement of the reference's V1 saturation model
// (internal/saturation/analyzer.go, constants.go) and of the live GPU-count
// limiter (internal/engines/pipeline/{default_limiter,type_inventory,
// greedy_saturation_algorithm}.go) over the SoA of include/wva_b200.h.
//
// Canonical order: the reference accumulates per-variant results in Go-map order
// (analyzer.go:86-94, random); here variants are visited in ascending index, and
// the shim indexes a model's variants by ascending VariantName, which also makes
// "tie -> alphabetically first/last name" (analyzer.go:391-393,421-423) an index
// comparison.  sort.Slice (greedy_saturation_algorithm.go:68) is unstable; here
// ties in (SpareCapacity, Cost) resolve to the lower decision index.
#pragma once
#include "../include/wva_b200.h"
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

namespace wva_oracle {

// internal/saturation/constants.go:8
static const int MinNonSaturatedReplicasForScaleDown = 2;

struct VariantAnalysis {  // interfaces.VariantSaturationAnalysis (saturation_analyzer.go:98-109)
  int ReplicaCount = 0, NonSaturatedCount = 0;
  double MaxKvCacheUsage = 0;
  int64_t MaxQueueLength = 0;
  double AvgSpareKvCapacity = 0, AvgSpareQueueLength = 0;
};

// analyzer.go:134-196 analyzeVariant
inline VariantAnalysis analyzeVariant(const wva_saturation_in& in, int64_t v, int64_t m, uint8_t* repSaturated) {
  VariantAnalysis a;
  int64_t lo = in.variant_replica_off[v], hi = in.variant_replica_off[v + 1];
  a.ReplicaCount = (int)(hi - lo);
  double kvThr = in.cfg_kv_threshold[m], qThr = in.cfg_queue_threshold[m];
  double totalSpareKv = 0, totalSpareQueue = 0;
  int nonSaturatedCount = 0;
  for (int64_t i = lo; i < hi; i++) {
    double kv = in.rep_kv[i];
    int64_t q = in.rep_queue[i];
    bool isSaturated = kv >= kvThr || (double)q >= qThr;  // :163-164
    if (repSaturated) repSaturated[i] = isSaturated ? 1 : 0;
    if (!isSaturated) {
      double spareKv = kvThr - kv;
      double spareQueue = qThr - (double)q;
      totalSpareKv += spareKv;
      totalSpareQueue += spareQueue;
      nonSaturatedCount++;
    }
    if (kv > a.MaxKvCacheUsage) a.MaxKvCacheUsage = kv;  // :179-184
    if (q > a.MaxQueueLength) a.MaxQueueLength = q;
  }
  a.NonSaturatedCount = nonSaturatedCount;
  if (nonSaturatedCount > 0) {  // :190-193
    a.AvgSpareKvCapacity = totalSpareKv / (double)nonSaturatedCount;
    a.AvgSpareQueueLength = totalSpareQueue / (double)nonSaturatedCount;
  }
  return a;
}

// analyzer.go:31-131 + 199-280 + 290-439 for every model of the batch.
inline void SaturationV1(const wva_saturation_in& in, const wva_saturation_out& out) {
  int64_t nUp = 0, nDown = 0, nTrans = 0, sumTargets = 0;
  for (int64_t m = 0; m < in.n_models; m++) {
    int64_t v0 = in.model_variant_off[m], v1 = in.model_variant_off[m + 1];
    // AnalyzeModelSaturation: only variants that HAVE metrics appear in VariantAnalyses
    // (variantMap is built from replicaMetrics, analyzer.go:62-78).
    double totalSpareKv = 0, totalSpareQueue = 0;
    int nonSaturatedCount = 0;
    int totalReplicas = 0;
    std::vector<VariantAnalysis> va((size_t)(v1 - v0));
    int nAnalysed = 0;
    for (int64_t v = v0; v < v1; v++) {
      VariantAnalysis a = analyzeVariant(in, v, m, out.rep_saturated);
      va[(size_t)(v - v0)] = a;
      if (a.ReplicaCount > 0) {
        nAnalysed++;
        nonSaturatedCount += a.NonSaturatedCount;                                     // :87-92
        totalSpareKv += a.AvgSpareKvCapacity * (double)a.NonSaturatedCount;
        totalSpareQueue += a.AvgSpareQueueLength * (double)a.NonSaturatedCount;
      }
      totalReplicas += a.ReplicaCount;
      if (out.var_replica_count) out.var_replica_count[v] = a.ReplicaCount;
      if (out.var_non_saturated) out.var_non_saturated[v] = a.NonSaturatedCount;
      if (out.var_max_kv) out.var_max_kv[v] = a.MaxKvCacheUsage;
      if (out.var_max_queue) out.var_max_queue[v] = a.MaxQueueLength;
      if (out.var_avg_spare_kv) out.var_avg_spare_kv[v] = a.AvgSpareKvCapacity;
      if (out.var_avg_spare_queue) out.var_avg_spare_queue[v] = a.AvgSpareQueueLength;
    }
    double avgSpareKv = 0, avgSpareQueue = 0;
    bool shouldScaleUp = false, scaleDownSafe = false, kvTrig = false, qTrig = false;
    if (totalReplicas > 0) {  // len(replicaMetrics)==0 => early return, all false (:39-50)
      if (nonSaturatedCount > 0) {  // :98-101
        avgSpareKv = totalSpareKv / (double)nonSaturatedCount;
        avgSpareQueue = totalSpareQueue / (double)nonSaturatedCount;
      }
      // shouldScaleUp :199-226
      kvTrig = avgSpareKv < in.cfg_kv_trigger[m];
      qTrig = avgSpareQueue < in.cfg_queue_trigger[m];
      shouldScaleUp = kvTrig || qTrig;
      // isScaleDownSafe :233-280
      if (nonSaturatedCount >= MinNonSaturatedReplicasForScaleDown) {
        double avgKvLoad = in.cfg_kv_threshold[m] - avgSpareKv;
        double avgQueueLoad = in.cfg_queue_threshold[m] - avgSpareQueue;
        int remainingCount = nonSaturatedCount - 1;
        double scaleFactor = (double)nonSaturatedCount / (double)remainingCount;
        double avgKvAfterRemoval = avgKvLoad * scaleFactor;
        double avgQueueAfterRemoval = avgQueueLoad * scaleFactor;
        double remainingSpareKv = in.cfg_kv_threshold[m] - avgKvAfterRemoval;
        double remainingSpareQueue = in.cfg_queue_threshold[m] - avgQueueAfterRemoval;
        bool kvSafe = remainingSpareKv >= in.cfg_kv_trigger[m];
        bool queueSafe = remainingSpareQueue >= in.cfg_queue_trigger[m];
        scaleDownSafe = kvSafe && queueSafe;
      }
    }
    if (out.mod_total_replicas) out.mod_total_replicas[m] = totalReplicas;
    if (out.mod_non_saturated) out.mod_non_saturated[m] = nonSaturatedCount;
    if (out.mod_avg_spare_kv) out.mod_avg_spare_kv[m] = avgSpareKv;
    if (out.mod_avg_spare_queue) out.mod_avg_spare_queue[m] = avgSpareQueue;

    // CalculateSaturationTargets :290-439
    std::vector<int> targets((size_t)(v1 - v0), 0);
    std::vector<char> hasTarget((size_t)(v1 - v0), 0);
    bool inTransition = false;
    auto cur = [&](int64_t v) { return (in.var_has_state && !in.var_has_state[v]) ? 0 : in.var_current[v]; };
    auto des = [&](int64_t v) { return (in.var_has_state && !in.var_has_state[v]) ? 0 : in.var_desired[v]; };
    auto pen = [&](int64_t v) { return (in.var_has_state && !in.var_has_state[v]) ? 0 : in.var_pending[v]; };
    if (nAnalysed == 0) {
      // nil safety :303-309: default = current replicas for every state
      for (int64_t v = v0; v < v1; v++) {
        if (in.var_has_state && !in.var_has_state[v]) continue;
        targets[(size_t)(v - v0)] = in.var_current[v];
        hasTarget[(size_t)(v - v0)] = 1;
      }
    } else {
      for (int64_t v = v0; v < v1; v++) {  // :322-341
        const VariantAnalysis& a = va[(size_t)(v - v0)];
        if (a.ReplicaCount == 0) continue;  // not in VariantAnalyses
        if (des(v) != 0 && des(v) != cur(v)) inTransition = true;
        if (a.ReplicaCount != cur(v)) inTransition = true;
      }
      for (int64_t v = v0; v < v1; v++) {  // :346-367
        const VariantAnalysis& a = va[(size_t)(v - v0)];
        if (a.ReplicaCount == 0) continue;
        hasTarget[(size_t)(v - v0)] = 1;
        if (inTransition) {
          if (des(v) != 0 && des(v) != cur(v)) targets[(size_t)(v - v0)] = des(v);
          else targets[(size_t)(v - v0)] = cur(v);
        } else {
          targets[(size_t)(v - v0)] = a.ReplicaCount;
        }
      }
      if (!inTransition) {
        if (shouldScaleUp) {  // :378-405
          int64_t cheapest = -1;
          for (int64_t v = v0; v < v1; v++) {
            if (va[(size_t)(v - v0)].ReplicaCount == 0) continue;
            if (pen(v) > 0) continue;
            if (cheapest < 0 || in.var_cost[v] < in.var_cost[cheapest]) cheapest = v;  // tie: lower index = name asc
          }
          if (cheapest >= 0) { targets[(size_t)(cheapest - v0)] += 1; nUp++; }
        } else if (scaleDownSafe) {  // :407-433
          int64_t expensive = -1;
          for (int64_t v = v0; v < v1; v++) {
            if (va[(size_t)(v - v0)].ReplicaCount == 0) continue;
            if (targets[(size_t)(v - v0)] <= 1) continue;
            if (expensive < 0 || in.var_cost[v] > in.var_cost[expensive] ||
                (in.var_cost[v] == in.var_cost[expensive] && v > expensive))  // tie: name desc
              expensive = v;
          }
          if (expensive >= 0) { targets[(size_t)(expensive - v0)] -= 1; nDown++; }
        }
      } else {
        nTrans++;
      }
    }
    uint8_t flags = 0;
    if (shouldScaleUp) flags |= WVA_SAT_SCALE_UP;
    if (scaleDownSafe) flags |= WVA_SAT_SCALE_DOWN_SAFE;
    if (inTransition) flags |= WVA_SAT_IN_TRANSITION;
    if (kvTrig) flags |= WVA_SAT_KV_TRIGGERED;
    if (qTrig) flags |= WVA_SAT_QUEUE_TRIGGERED;
    if (out.mod_flags) out.mod_flags[m] = flags;
    for (int64_t v = v0; v < v1; v++) {
      // variants absent from the targets map are reported as -1
      int t = hasTarget[(size_t)(v - v0)] ? targets[(size_t)(v - v0)] : -1;
      if (out.var_target) out.var_target[v] = t;
      if (t >= 0) sumTargets += t;
    }
  }
  if (out.partials) {
    out.partials[0] = nUp;
    out.partials[1] = nDown;
    out.partials[2] = nTrans;
    out.partials[3] = sumTargets;
  }
}

// DefaultLimiter.Limit (default_limiter.go:42-81) with TypeInventory
// (type_inventory.go:222-243) / typeAllocator.TryAllocate (:347-373) and
// GreedyBySaturation (greedy_saturation_algorithm.go:34-108).
inline void Limit(int64_t D, int T, const int32_t* accType, const int32_t* current, const int32_t* target,
                  const int32_t* gpusPerReplica, const double* spare, const double* cost,
                  const int32_t* typeLimit, int32_t* outTarget, int32_t* outGpus, uint8_t* outLimited) {
  // calculateUsedGPUs default_limiter.go:72-81 (raw GPUsPerReplica, not defaulted)
  std::vector<int64_t> used((size_t)T, 0);
  for (int64_t d = 0; d < D; d++) {
    if (accType[d] < 0) continue;
    used[(size_t)accType[d]] += (int64_t)current[d] * gpusPerReplica[d];
  }
  // CreateAllocator type_inventory.go:222-243
  std::vector<int64_t> remaining((size_t)T);
  for (int t = 0; t < T; t++) remaining[(size_t)t] = std::max<int64_t>((int64_t)typeLimit[t] - used[(size_t)t], 0);
  for (int64_t d = 0; d < D; d++) { outTarget[d] = target[d]; outGpus[d] = 0; outLimited[d] = 0; }
  // filterScaleUpCandidates :50-58, sortByPriority :65-75
  std::vector<int64_t> c;
  for (int64_t d = 0; d < D; d++) if (target[d] > current[d]) c.push_back(d);
  std::stable_sort(c.begin(), c.end(), [&](int64_t i, int64_t j) {
    if (spare[i] != spare[j]) return spare[i] < spare[j];
    return cost[i] < cost[j];
  });
  for (int64_t d : c) {  // allocateForDecision :79-108
    int replicasNeeded = target[d] - current[d];
    if (replicasNeeded <= 0) continue;
    int gpr = gpusPerReplica[d];
    if (gpr <= 0) gpr = 1;
    int64_t gpusRequested = (int64_t)replicasNeeded * gpr;
    int64_t gpusAllocated = 0;
    if (gpusRequested > 0 && accType[d] >= 0) {  // TryAllocate :347-373 ("" -> error, 0 allocated)
      int64_t avail = remaining[(size_t)accType[d]];
      if (avail > 0) {
        gpusAllocated = std::min(gpusRequested, avail);
        remaining[(size_t)accType[d]] -= gpusAllocated;
      }
    }
    int64_t replicasAllocated = gpusAllocated / gpr;
    outGpus[d] = (int32_t)(replicasAllocated * gpr);
    outTarget[d] = current[d] + (int32_t)replicasAllocated;
    if (replicasAllocated < replicasNeeded) outLimited[d] = 1;
  }
}

}  // namespace wva_oracle
