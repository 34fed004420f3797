// run_plan.hpp -- host-side planning of the run kernel's work (pure function, no CUDA calls): which records of which
// (sub-stripe, item) run bucket every launch ("visit" = sub-stripe x round x item block) walks, cut into runs.
//   * a bucket is spread over as many of the `rounds` interleaved passes as it has runs of >= 2 * MIN_RUN records for
//     (frequently rated items: all of them); a small bucket is walked whole in one pass -- which one is hashed from
//     (sub-stripe, item), so the passes stay balanced;
//   * a pass's slice is cut into ceil(n / chunk) equal runs; a run's merge weight is min(1, boost / runs of its slice), i.e.
//     the runs of one item in one launch are averaged, slightly over-relaxed (kernels_hot.cu; merge_weight below);
//   * inside a visit the runs are ordered longest first (stable), so the launch's tail is made of short runs and the
//     runs a warp walks side by side have about the same length.
// Exposed for CPU tests through mfsgd_plan_runs (include/mfsgd.h).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <new>
#include <thread>
#include <vector>

#include "../../include/mfsgd.h"
#include "kernels.cuh"

namespace mfsgd {

static const int MIN_RUN = 16;   // shortest run worth a sub-warp of the run kernel (one q_i load + merge per run)

// ---- blocking of a ring member's work (pure functions of the configuration) -----------------------------------------
// Sub-stripes of P per ring member (mu) and item sub-shards per member (mi). Auto mu keeps one P sub-stripe plus the
// held Q shard group resident in L2 (measured on the Netflix-shaped workload: 61 MB sub-stripes beat 35 MB and 82 MB
// ones); auto mi = 2 where the rotation is pipelined (a multi-process ring sends one slice of Q while the other trains;
// the two slices' launches run on two stream lanes and fill each other's tails, engine.cu lane mode).
struct Blocking {
    int mu, mi;
};
inline Blocking plan_blocking(int n_users, int n_items, int k, int G, int mode, int stripes_per_gpu, int shards_per_gpu,
                              bool multi_process, double l2_bytes) {
    Blocking b;
    b.mi = shards_per_gpu > 0 ? shards_per_gpu : (multi_process && G >= 2 && mode == MFSGD_MODE_DSGD ? 2 : 1);
    if (mode == MFSGD_MODE_DETERMINISTIC) {
        b.mu = 1;
    } else if (stripes_per_gpu > 0) {
        b.mu = stripes_per_gpu;
    } else {
        const double p_bytes = (double)n_users * k * 4.0 / G;
        const double q_bytes = (double)n_items * k * 4.0 / G;
        int mu = 1;
        if (l2_bytes > 0 && p_bytes + q_bytes > 0.6 * l2_bytes) {
            const double budget = std::max(0.6 * l2_bytes - q_bytes, 0.1 * l2_bytes);
            mu = (int)std::ceil(p_bytes / budget);
        }
        b.mu = std::min(std::max(mu, 1), 256);
    }
    return b;
}

// Interleaved passes per sub-epoch. A sub-epoch sweeps its mu sub-stripes `rounds` times, a slice of each bucket per
// visit, so Q sees every user stripe many times per epoch (a plain stripe-after-stripe order biases the item factors
// toward the last stripe and costs ~1 % RMSE at equal epochs) while each visit is still long enough (>= ~16 touches
// per P row) to keep the sub-stripe L2-resident. The run path averages the runs of an item that share a launch, so an
// item's factor makes sequential progress only from launch to launch: keep >= 8 launches per epoch (G * mu * rounds)
// while a launch still holds >= 16 K records (a 1.8 M-rating set trained with one launch per epoch ends 1 % above the
// oracle's RMSE after 8 epochs; with 8 it lands on it).
inline int plan_rounds(int rounds_cfg, int mode, int G, int mu, int64_t member_records, int64_t member_users) {
    if (rounds_cfg > 0) return rounds_cfg;
    if (mode == MFSGD_MODE_DETERMINISTIC) return 1;
    int rounds = 1;
    if (mu > 1) {
        const double block_recs = (double)member_records / ((double)mu * G);
        const double stripe_rows = std::max(1.0, (double)member_users / mu);
        rounds = (int)std::min(4.0, std::max(1.0, std::floor(block_recs / (16.0 * stripe_rows))));
    }
    while (G * mu * rounds < 8 && (double)member_records / ((double)mu * rounds * 2) >= 16384.0) rounds *= 2;
    return rounds;
}

// Longest run. A run is walked by one sub-warp, one rating after another, and the runs of an item that share a launch
// all start from the same q_i and are merged by a weighted average (plan_runs): every split of an item's slice costs
// sequential progress on that item. Round 2 (tools/run_sim.py, signal-dominant sets -- on round 1's noise-dominant sets the
// effect was invisible): runs of <= 256 trail the sequential oracle by 24 % / 9 % / 3.7 % in held-out RMSE after epochs
// 1 / 2 / 3 of a Netflix-shaped set, runs of <= 1024 by 1.4 % / 0.1 % / 0.2 %; every cleverer merge that was tried
// (summing, direction-split weights, exchanging q_i inside the run) diverges on the stiff common direction that plain
// matrix factorisation without biases has. So runs are as long as the launch can balance: the launch hands its runs out
// longest first, so it lasts about max(records / resident sub-warps, longest run) -- the longest run may be about the
// per-sub-warp share of the launch.
// 256 <= run <= 1024 (from 128 for laned launches), multiple of 32: small launches leave sub-warps idle rather than cut their items into pieces
// (they are latency-bound anyway: an ML-100K-shaped epoch takes 0.35 ms either way).
inline int plan_run_length(int hot_chunk_cfg, int lanes, int mu, int rounds, int IB, int64_t run_records, int resident_ctas,
                           int runs_per_warp) {
    if (hot_chunk_cfg > 0) return hot_chunk_cfg;
    // Laned launches (the item sub-shards of a one-process-per-GPU ring on two stream lanes) share the machine, so a lane's launch
    // lasts about twice its per-sub-warp share and a run may be that long (`lanes` launches' worth) -- as long as that still gives
    // runs above the floor of 256. Below it the launches are too small for that: measured on 8 B200s (Netflix-shaped, 0.7 M ratings
    // per launch) runs of 256 / 192 / 128 take 1.50 / 1.32 / 1.16 ms per epoch = 5.5x / 6.3x / 7.2x one GPU, because at 256 a launch
    // offers too few runs to fill its half of the machine twice over (profiles/r02_experiments.md section 11). Such launches get
    // runs of ONE share, down to 128.
    const double per_launch = (double)run_records / ((double)mu * rounds * IB);
    const double share = per_launch / (1.25 * resident_ctas * 8.0 * runs_per_warp);
    const double laned_want = share * (lanes > 1 ? lanes : 1);
    if (lanes > 1 && laned_want <= 256.0) return (int)std::max(128.0, std::ceil(share / 32.0) * 32.0);
    return (int)std::min(1024.0, std::max(256.0, std::ceil(laned_want / 32.0) * 32.0));
}

// Merge weight of a run whose slice of its item's bucket was cut into `pieces` runs for one launch: min(1, boost / pieces).
// boost = 1 is plain model averaging; the default 1.25 over-relaxes the average a little (the runs of a slice each saw only
// 1 / pieces of its ratings, so their mean under-shoots what a sequential walk of the whole slice reaches; summing
// them -- boost = pieces -- diverges). Chosen on tools/run_sim.py: held-out RMSE within +-0.6 % of the sequential oracle
// at every epoch of the signal-dominant Netflix-shaped set (plain averaging ends 0.5 % below it, boost 1.9 3 % above).
static const float DEFAULT_MERGE_BOOST = 1.25f;
inline float merge_weight(int64_t pieces, float boost) {
    if (pieces <= 1) return 1.0f;
    const float w = (boost > 1.0f ? boost : 1.0f) / (float)pieces;
    return w > 1.0f ? 1.0f : w;
}

struct RunPlanArgs {
    const int64_t* block_off;      // offsets of the member's buckets: mu * IB cold blocks, then mu * H run buckets (+1)
    size_t n_blocks;               // mu * (IB + H)
    int mu, H, IB, rounds, chunk;  // sub-stripes, run items, item blocks, passes per sub-epoch, longest run
    int member;                    // ring member g (keys the buckets' per-epoch permutations)
    uint64_t seed;
    float boost;                   // merge over-relaxation (merge_weight)
    const int32_t* hot_block_lo;   // IB + 1: run items [hot_block_lo[b], hot_block_lo[b+1]) lie in item block b
    const int32_t* hot_items;      // H global item ids, ascending
};

// units: all runs, visit after visit; visit_units[(sa * rounds + rnd) * IB + ib] = first run of that visit (+1 entry: total)
// The plan is on the critical path of every load (16-27 ms of a 226 ms end-to-end factorize call on the Netflix-shaped set:
// 71 K buckets, 224 K runs), so it is a two-pass counting sort with no intermediate lists, cut over host threads when it is
// large: thread t takes the t-th range of run items (every sub-stripe), pass 1 counts its runs per (visit, run length), a
// prefix sum in (visit, longest first, thread) order gives every thread its slots, pass 2 enumerates the same runs again and
// writes each into its slot. Thread order is item order, so the result is the serial stable sort, bit for bit, whatever the
// thread count (tests/test_run_plan_cpu.py).
inline void plan_runs(const RunPlanArgs& a, std::vector<HotUnit>& units, std::vector<int>& visit_units, int threads = 0) {
    units.clear();
    const size_t n_visits = (size_t)a.mu * a.rounds * a.IB;
    visit_units.assign(n_visits + 1, 0);
    if (a.H <= 0 || a.chunk <= 0 || a.mu <= 0 || a.rounds <= 0 || a.IB <= 0) return;
    const size_t hot_base = (size_t)a.mu * a.IB;
    const size_t keys = (size_t)a.chunk + 1;                 // key = chunk - run length: 0 = longest
    int T = threads;
    if (T <= 0) {
        T = 1;
        if ((size_t)a.mu * (size_t)a.H >= 16384) T = (int)std::min<unsigned>(8u, std::max<unsigned>(1u, std::thread::hardware_concurrency()));
    }
    T = std::max(1, std::min(T, a.H));
    while (T > 1 && (size_t)T * n_visits * keys > ((size_t)1 << 24)) T--;      // slot tables: at most 128 MB
    std::vector<uint64_t> slot((size_t)T * n_visits * keys, 0);               // pass 1: counts; after the prefix sum: next free slot

    // the runs of run items [H t / T, H (t + 1) / T), every sub-stripe, in the order (sub-stripe, item block, item, slice, piece)
    auto enumerate = [&](int t, auto&& emit) {
        const int hx_lo = (int)((int64_t)a.H * t / T), hx_hi = (int)((int64_t)a.H * (t + 1) / T);
        for (int sa = 0; sa < a.mu; sa++)
            for (int ib = 0; ib < a.IB; ib++) {
                const int lo_x = std::max(hx_lo, (int)a.hot_block_lo[ib]), hi_x = std::min(hx_hi, (int)a.hot_block_lo[ib + 1]);
                for (int hx = lo_x; hx < hi_x; hx++) {
                    const size_t blk = hot_base + (size_t)sa * a.H + (size_t)hx;
                    const int64_t bn = a.block_off[blk + 1] - a.block_off[blk];
                    if (bn <= 0) continue;
                    const int spread = (int)std::min<int64_t>(a.rounds, std::max<int64_t>(1, bn / (2 * MIN_RUN)));
                    const int first = (int)(hash64(a.seed, 11, ((uint64_t)sa << 32) | (uint64_t)(uint32_t)a.hot_items[hx]) % (uint64_t)a.rounds);
                    for (int sl = 0; sl < spread; sl++) {
                        const int rnd = (first + sl * a.rounds / spread) % a.rounds;     // distinct for distinct sl (spread <= rounds)
                        const int64_t lo = a.block_off[blk] + bn * sl / spread, hi = a.block_off[blk] + bn * (sl + 1) / spread;
                        const int64_t n = hi - lo;
                        if (n <= 0) continue;
                        const int64_t pieces = (n + a.chunk - 1) / a.chunk;
                        const size_t visit = ((size_t)sa * a.rounds + rnd) * a.IB + ib;
                        const float w = merge_weight(pieces, a.boost);
                        int64_t at = lo;
                        for (int64_t pc = 0; pc < pieces; pc++) {
                            const int64_t next = lo + n * (pc + 1) / pieces;
                            emit(visit, blk, bn, hx, at, (int32_t)(next - at), w);
                            at = next;
                        }
                    }
                }
            }
    };
    auto fan_out = [&](auto&& job) {     // job(t) for t in [0, T) on T threads (the caller's is one of them)
        std::vector<std::thread> pool;
        try {
            for (int t = 1; t < T; t++) pool.emplace_back(job, t);
        } catch (...) {                  // no thread to be had: the caller does the rest itself
        }
        const int started = (int)pool.size() + 1;
        job(0);
        for (std::thread& th : pool) th.join();
        for (int t = started; t < T; t++) job(t);
    };
    fan_out([&](int t) {
        uint64_t* mine = slot.data() + (size_t)t * n_visits * keys;
        enumerate(t, [&](size_t visit, size_t, int64_t, int, int64_t, int32_t count, float) { mine[visit * keys + (size_t)(a.chunk - count)]++; });
    });
    uint64_t run = 0;
    for (size_t v = 0; v < n_visits; v++) {
        visit_units[v] = (int)run;
        for (size_t key = 0; key < keys; key++)
            for (int t = 0; t < T; t++) {
                uint64_t& c = slot[((size_t)t * n_visits + v) * keys + key];
                const uint64_t cnt = c;
                c = run;
                run += cnt;
            }
    }
    visit_units[n_visits] = (int)run;
    units.resize((size_t)run);
    fan_out([&](int t) {
        uint64_t* mine = slot.data() + (size_t)t * n_visits * keys;
        enumerate(t, [&](size_t visit, size_t blk, int64_t bn, int hx, int64_t start, int32_t count, float w) {
            HotUnit u{};
            u.bstart = a.block_off[blk];
            u.bn = (int32_t)bn;
            u.bid = (uint32_t)((size_t)a.member * a.n_blocks + blk);
            u.start = start;
            u.count = count;
            u.item = a.hot_items[hx];
            u.weight = w;
            units[(size_t)mine[visit * keys + (size_t)(a.chunk - count)]++] = u;
        });
    });
}

}  // namespace mfsgd
