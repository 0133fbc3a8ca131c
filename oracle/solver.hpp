// ORACLE — TEST INFRASTRUCTURE ONLY (see analyzer.hpp header).
// CPU restatement of the reference's pkg/solver (solver.go, greedy.go) and of
// System.AllocateByType (pkg/core/system.go:271-299).
//
// Canonical order: wherever the reference iterates a Go map (random order) or
// calls the unstable slices.SortFunc, this restatement iterates ascending index
// and uses a STABLE sort, i.e. ties resolve to the lower canonical index.  The
// reinsertion rule of allocate() (slices.BinarySearchFunc -> insert BEFORE equal
// elements, greedy.go:161-162) is deterministic in the reference and is kept
// literally.
//
// Citations are relative to /root/reference/pkg/solver/ unless noted.
#pragma once
#include "core.hpp"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <map>
#include <vector>

namespace wva_oracle {

// Go cmp.Compare for floats (NaN sorts first)
inline int cmpCompare(float x, float y) {
  bool xn = std::isnan(x), yn = std::isnan(y);
  if (xn) return yn ? 0 : -1;
  if (yn) return +1;
  if (x < y) return -1;
  if (x > y) return +1;
  return 0;
}
inline int cmpCompareInt(int x, int y) { return x < y ? -1 : (x > y ? +1 : 0); }

struct Solution {
  std::vector<Allocation> alloc;      // [S] Server.Allocation() (state NONE = nil)
  std::vector<int64_t> typeCount;     // [T]
  std::vector<float> typeCostF32;     // [T] float32 running sum in canonical order (as the reference)
  std::vector<double> typeCost;       // [T] same sum in float64
};

// solver.go:63-79
inline void SolveUnlimited(const wva_system& s, const std::vector<Allocation>& cand, std::vector<Allocation>& sol) {
  const int S = s.n_servers, A = s.n_acc;
  sol.assign((size_t)S, Allocation{});
  for (int srv = 0; srv < S; srv++) {
    float minVal = FLT_MAX;
    const Allocation* minAlloc = nullptr;
    for (int g = 0; g < A; g++) {
      const Allocation& a = cand[(size_t)srv * A + g];
      if (a.state == WVA_ALLOC_NONE) continue;
      if (a.value < minVal) { minVal = a.value; minAlloc = &a; }
    }
    if (minAlloc) sol[srv] = *minAlloc;
  }
}

// greedy.go:16-22
struct serverEntry {
  int server;
  int priority;
  int curIndex;
  std::vector<Allocation> allocations;  // sorted by value
  float delta;
};

// greedy.go:76-87
inline int orderFunc(const serverEntry* a, const serverEntry* b) {
  if (a->priority == b->priority) {
    if (a->delta == b->delta)
      return cmpCompare(b->allocations[b->curIndex].value, a->allocations[a->curIndex].value);
    return cmpCompare(b->delta, a->delta);
  }
  return cmpCompareInt(a->priority, b->priority);
}

// unitsPerReplica := model.NumInstances(gName) * acc.Spec().Multiplicity (greedy.go:139)
inline int unitsPerReplica(const wva_system& s, int srv, int acc) {
  return NumInstances(s, s.srv_model[srv], acc) * s.acc_multiplicity[acc];
}

// greedy.go:107-166
inline std::vector<serverEntry*> allocate(const wva_system& s, std::vector<serverEntry*> entries,
                                          std::vector<int64_t>& available, std::vector<Allocation>& sol) {
  std::vector<serverEntry*> unallocated;
  size_t head = 0;  // entries = entries[1:] is a slice re-header; modelled by a moving head
  while (head < entries.size()) {
    serverEntry* top = entries[head];
    head++;
    if (top->allocations.empty()) continue;
    int srv = top->server;
    if (s.srv_model[srv] < 0) continue;                       // model == nil
    Allocation& alloc = top->allocations[top->curIndex];
    if (alloc.state != WVA_ALLOC_ACC) continue;                // accelerator "" -> GetAccelerator nil (greedy.go:135-137)
    int gName = alloc.acc;
    int tName = s.acc_type[gName];
    int64_t count = (int64_t)alloc.numReplicas * unitsPerReplica(s, srv, gName);
    if (available[tName] >= count) {                          // greedy.go:143-145
      available[tName] -= count;
      sol[srv] = alloc;
    } else {
      top->curIndex++;
      if ((size_t)top->curIndex + 1 < top->allocations.size()) {
        top->delta = top->allocations[top->curIndex + 1].value - top->allocations[top->curIndex].value;
      } else if ((size_t)top->curIndex == top->allocations.size()) {
        unallocated.push_back(top);
        continue;
      } else {
        top->delta = FLT_MAX;
      }
      // slices.BinarySearchFunc: smallest i in the remaining list with cmp(entries[i], top) >= 0
      size_t lo = head, hi = entries.size();
      while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (orderFunc(entries[mid], top) < 0) lo = mid + 1;
        else hi = mid;
      }
      entries.insert(entries.begin() + (ptrdiff_t)lo, top);   // slices.Insert
    }
  }
  return unallocated;
}

// greedy.go:321-341
inline std::vector<std::vector<serverEntry*>> makePriorityGroups(const std::vector<serverEntry*>& e) {
  std::vector<std::vector<serverEntry*>> groups;
  size_t index = 0, n = e.size();
  while (index < n) {
    std::vector<serverEntry*> group;
    group.push_back(e[index]);
    int groupPriority = e[index]->priority;
    index++;
    while (index < n && e[index]->priority == groupPriority) { group.push_back(e[index]); index++; }
    groups.push_back(std::move(group));
  }
  return groups;
}

// greedy.go:194-223
inline void allocateMaximally(const wva_system& s, const std::vector<serverEntry*>& entries,
                              std::vector<int64_t>& available, std::vector<Allocation>& sol) {
  for (serverEntry* entry : entries) {
    for (Allocation& alloc : entry->allocations) {
      if (alloc.state != WVA_ALLOC_ACC) continue;            // acc == nil
      int srv = entry->server;
      if (s.srv_model[srv] < 0) continue;
      int acc = alloc.acc;
      int upr = unitsPerReplica(s, srv, acc);
      if (upr > 0) {
        int t = s.acc_type[acc];
        int64_t maxReplicas = available[t] / upr;
        maxReplicas = std::min<int64_t>(maxReplicas, alloc.numReplicas);
        if (maxReplicas > 0) {
          int curNumReplicas = alloc.numReplicas;
          float factor = (float)maxReplicas / (float)curNumReplicas;
          alloc.cost = alloc.cost * factor;
          alloc.value = alloc.value * factor;
          alloc.numReplicas = (int)maxReplicas;
          sol[srv] = alloc;
          available[t] -= maxReplicas * upr;
          break;
        }
      }
    }
  }
}

// greedy.go:225-316
inline void allocateEqually(const wva_system& s, const std::vector<serverEntry*>& entries,
                            std::vector<int64_t>& available, std::vector<Allocation>& sol) {
  struct ticket_t {
    serverEntry* entry;
    bool active = false;
    int accType = -1;
    int unitsPerReplica = 0;
    int numReplicas = 0;
    Allocation* finalAlloc = nullptr;
    bool live = false;       // present in `tickets`
    bool allocated = false;  // present in `allocatedTickets`
  };
  // tickets keyed by server name; one entry per server here
  std::map<int, ticket_t> tickets;
  size_t nLive = 0;
  for (serverEntry* e : entries) {
    if (s.srv_model[e->server] < 0) continue;
    ticket_t t;
    t.entry = e;
    t.live = true;
    if (!tickets.count(e->server)) nLive++;
    tickets[e->server] = t;
  }
  while (nLive > 0) {
    for (serverEntry* e : entries) {
      auto it = tickets.find(e->server);
      if (it == tickets.end() || !it->second.live) continue;
      ticket_t& ticket = it->second;
      if (!ticket.active) {
        for (Allocation& alloc : e->allocations) {
          if (alloc.state != WVA_ALLOC_ACC) continue;
          int acc = alloc.acc;
          int upr = unitsPerReplica(s, e->server, acc);
          if (upr > 0 && available[s.acc_type[acc]] >= upr) {
            ticket.active = true;
            ticket.accType = s.acc_type[acc];
            ticket.unitsPerReplica = upr;
            ticket.finalAlloc = &alloc;
            break;
          }
        }
        if (!ticket.active) { ticket.live = false; nLive--; continue; }
      }
      int64_t replicasAvailable = available[ticket.accType] / ticket.unitsPerReplica;
      int64_t replicasAllocatable = std::min<int64_t>(replicasAvailable, ticket.finalAlloc->numReplicas);
      if (replicasAllocatable > 0) {
        ticket.numReplicas++;
        available[ticket.accType] -= ticket.unitsPerReplica;
        ticket.allocated = true;
      } else {
        ticket.live = false;
        nLive--;
      }
    }
  }
  for (auto& kv : tickets) {
    ticket_t& ticket = kv.second;
    if (!ticket.allocated) continue;
    Allocation& alloc = *ticket.finalAlloc;
    int numReplicas = ticket.numReplicas;
    int curNumReplicas = alloc.numReplicas;
    float factor = (float)numReplicas / (float)curNumReplicas;
    alloc.cost = alloc.cost * factor;
    alloc.value = alloc.value * factor;
    alloc.numReplicas = numReplicas;
    sol[ticket.entry->server] = alloc;
  }
}

// greedy.go:169-192
inline void bestEffort(const wva_system& s, const std::vector<serverEntry*>& unallocated,
                       std::vector<int64_t>& available, int policy, std::vector<Allocation>& sol) {
  switch (policy) {
    case WVA_POLICY_PRIORITY_EXHAUSTIVE:
      allocateMaximally(s, unallocated, available, sol);
      break;
    case WVA_POLICY_PRIORITY_ROUND_ROBIN: {
      auto groups = makePriorityGroups(unallocated);
      for (auto& g : groups) allocateEqually(s, g, available, sol);
      break;
    }
    case WVA_POLICY_ROUND_ROBIN:
      allocateEqually(s, unallocated, available, sol);
      break;
    default:
      break;
  }
}

// greedy.go:35-105
inline void SolveGreedy(const wva_system& s, const std::vector<Allocation>& cand, std::vector<Allocation>& sol) {
  const int S = s.n_servers, A = s.n_acc, T = s.n_types;
  sol.assign((size_t)S, Allocation{});
  std::vector<int64_t> available((size_t)T);
  for (int t = 0; t < T; t++) available[t] = s.type_count[t];
  std::vector<serverEntry> store;
  store.reserve((size_t)S);
  for (int srv = 0; srv < S; srv++) {
    serverEntry e;
    e.server = srv;
    e.priority = s.srv_priority[srv];
    e.curIndex = 0;
    e.delta = 0;
    for (int g = 0; g < A; g++) {
      const Allocation& a = cand[(size_t)srv * A + g];
      if (a.state != WVA_ALLOC_NONE) e.allocations.push_back(a);
    }
    if (e.allocations.empty()) continue;
    std::stable_sort(e.allocations.begin(), e.allocations.end(),
                     [](const Allocation& a, const Allocation& b) { return cmpCompare(a.value, b.value) < 0; });
    if (e.allocations.size() > 1) e.delta = e.allocations[1].value - e.allocations[0].value;
    else e.delta = FLT_MAX;
    store.push_back(std::move(e));
  }
  std::vector<serverEntry*> entries;
  entries.reserve(store.size());
  for (auto& e : store) entries.push_back(&e);
  std::stable_sort(entries.begin(), entries.end(),
                   [](const serverEntry* a, const serverEntry* b) { return orderFunc(a, b) < 0; });
  if (s.delayed_best_effort) {
    auto un = allocate(s, entries, available, sol);
    bestEffort(s, un, available, s.saturation_policy, sol);
  } else {
    auto groups = makePriorityGroups(entries);
    for (auto& g : groups) {
      auto un = allocate(s, g, available, sol);
      bestEffort(s, un, available, s.saturation_policy, sol);
    }
  }
}

// pkg/core/system.go:271-299
inline void AllocateByType(const wva_system& s, Solution& out) {
  const int T = s.n_types;
  out.typeCount.assign((size_t)T, 0);
  out.typeCostF32.assign((size_t)T, 0.0f);
  out.typeCost.assign((size_t)T, 0.0);
  for (int srv = 0; srv < s.n_servers; srv++) {
    const Allocation& a = out.alloc[srv];
    if (a.state != WVA_ALLOC_ACC) continue;  // nil, or accelerator "" (acc == nil)
    int model = s.srv_model[srv];
    if (model < 0) continue;
    int t = s.acc_type[a.acc];
    out.typeCount[t] += (int64_t)a.numReplicas * NumInstances(s, model, a.acc) * s.acc_multiplicity[a.acc];
    out.typeCostF32[t] += a.cost;
    out.typeCost[t] += (double)a.cost;
  }
}

// solver.go:32-60 + pkg/manager/manager.go:21-27
inline void ManagerOptimize(const wva_system& s, const std::vector<Allocation>& cand, Solution& out) {
  if (s.unlimited) SolveUnlimited(s, cand, out.alloc);
  else SolveGreedy(s, cand, out.alloc);
  AllocateByType(s, out);
}

}  // namespace wva_oracle
