// CPU unit test of the batching-rounds barrier (tensornetworks.jl_b200/csrc/tn_rounds.h), the synchronisation logic behind
// TN_QJMC_BATCH=1.  Built and run by tests/test_rounds_cpu.py (g++ -pthread) under a time limit: a dead-lock fails the test.
#include "../../tensornetworks.jl_b200/csrc/tn_rounds.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <random>
#include <thread>

struct Req { int owner; int seq; int executed = 0; unsigned long long round = 0; };

static int fail(const char* msg) { std::printf("FAIL: %s\n", msg); return 1; }

// every thread submits counts[i] requests with random pauses, then leaves
static int scenario(const std::vector<int>& counts, int throw_at_round, unsigned seed) {
  const int n = (int)counts.size();
  tn::Rounds<Req> rounds(n);
  std::atomic<long long> executed{0};
  std::atomic<int> max_group{0}, exceptions{0};
  std::vector<std::vector<Req>> reqs(n);
  for (int i = 0; i < n; ++i) { reqs[i].resize(counts[i]); for (int k = 0; k < counts[i]; ++k) { reqs[i][k].owner = i; reqs[i][k].seq = k; } }
  std::atomic<int> in_exec{0};
  bool overlap = false, dup_owner = false;
  auto exec = [&](std::vector<Req*>& pend) {
    if (in_exec.fetch_add(1) != 0) overlap = true;                    // exec must never run concurrently
    if (throw_at_round >= 0 && (int)rounds.rounds() == throw_at_round) { in_exec.fetch_sub(1); throw std::runtime_error("boom"); }
    std::vector<int> seen(n, 0);
    for (Req* r : pend) { r->executed++; r->round = rounds.rounds(); if (seen[r->owner]++) dup_owner = true; }
    executed += (long long)pend.size();
    int g = (int)pend.size(), m = max_group.load();
    while (g > m && !max_group.compare_exchange_weak(m, g)) {}
    in_exec.fetch_sub(1);
  };
  std::vector<std::thread> pool;
  for (int i = 0; i < n; ++i)
    pool.emplace_back([&, i] {
      std::mt19937 rng(seed * 7919u + i);
      try {
        for (int k = 0; k < counts[i]; ++k) {
          if (rng() % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 200));
          rounds.submit(&reqs[i][k], exec);
          if (!reqs[i][k].executed) { std::printf("FAIL: submit returned before its request ran\n"); std::abort(); }
        }
      } catch (const std::runtime_error&) { exceptions++; }
      rounds.leave(exec);
    });
  for (auto& t : pool) t.join();
  if (overlap) return fail("exec ran concurrently");
  if (dup_owner) return fail("two requests of one participant in the same round");
  long long total = 0; for (int c : counts) total += c;
  if (throw_at_round < 0) {
    if (exceptions != 0) return fail("unexpected exception");
    if (executed != total) return fail("not every request was executed");
    for (auto& v : reqs) for (auto& r : v) if (r.executed != 1) return fail("a request ran more or less than once");
    int longest = 0; for (int c : counts) longest = std::max(longest, c);
    if ((long long)rounds.rounds() < longest) return fail("fewer rounds than the longest sequence");
    if (max_group > n) return fail("group larger than the number of participants");
    if (rounds.requests() != total) return fail("request counter");
  } else {
    if (!rounds.failed()) return fail("the failure was not recorded");
    int expect = 0; for (int c : counts) if (c > throw_at_round) expect++;    // everybody still submitting at / after the failing round throws
    if (exceptions < 1 || exceptions > n) return fail("exception count");
    (void)expect;
  }
  return 0;
}

int main() {
  int bad = 0;
  for (unsigned seed = 0; seed < 20; ++seed) {
    bad += scenario({50, 57, 64, 71, 78, 85, 92, 99}, -1, seed);                 // staggered exits
    bad += scenario({40, 40, 40, 40}, -1, seed);                                // perfect lockstep
    bad += scenario({0, 30, 0, 30, 5}, -1, seed);                               // participants that leave at once
    bad += scenario({25}, -1, seed);                                            // single participant
    bad += scenario({60, 60, 60, 60, 60, 60}, 10, seed);                        // exec throws in round 10: nobody may hang
    bad += scenario({3, 60, 60, 12}, 7, seed);
  }
  if (bad) { std::printf("%d scenario(s) failed\n", bad); return 1; }
  std::printf("rounds ok\n");
  return 0;
}
