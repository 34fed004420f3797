// run_plan.hpp -- host-side planning of the run kernel's work (pure function, no CUDA calls): which records of which
// (sub-stripe, item) run bucket every launch ("visit" = sub-stripe x round x item block) walks, cut into runs.
//   * a bucket is spread over as many of the `rounds` interleaved passes as it has runs of >= 2 * MIN_RUN records for
//     (frequently rated items: all of them); a small bucket is walked whole in one pass -- which one is hashed from
//     (sub-stripe, item), so the passes stay balanced;
//   * a pass's slice is cut into ceil(n / chunk) equal runs; a run's merge weight is 1 / (runs of its slice), i.e.
//     the runs of one item in one launch are averaged (kernels_hot.cu);
//   * inside a visit the runs are ordered longest first (stable), so the launch's tail is made of short runs and the
//     runs a warp walks side by side have about the same length.
// Exposed for CPU tests through mfsgd_plan_runs (include/mfsgd.h).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "kernels.cuh"

namespace mfsgd {

static const int MIN_RUN = 16;   // shortest run worth a sub-warp of the run kernel (one q_i load + merge per run)

struct RunPlanArgs {
    const int64_t* block_off;      // offsets of the member's buckets: mu * IB cold blocks, then mu * H run buckets (+1)
    size_t n_blocks;               // mu * (IB + H)
    int mu, H, IB, rounds, chunk;  // sub-stripes, run items, item blocks, passes per sub-epoch, longest run
    int member;                    // ring member g (keys the buckets' per-epoch permutations)
    uint64_t seed;
    const int32_t* hot_block_lo;   // IB + 1: run items [hot_block_lo[b], hot_block_lo[b+1]) lie in item block b
    const int32_t* hot_items;      // H global item ids, ascending
};

// units: all runs, visit after visit; visit_units[(sa * rounds + rnd) * IB + ib] = first run of that visit (+1 entry: total)
inline void plan_runs(const RunPlanArgs& a, std::vector<HotUnit>& units, std::vector<int>& visit_units) {
    units.clear();
    visit_units.assign((size_t)a.mu * a.rounds * a.IB + 1, 0);
    const size_t hot_base = (size_t)a.mu * a.IB;
    if (a.H > 0 && a.chunk > 0)
        units.reserve((size_t)((a.block_off[a.n_blocks] - a.block_off[hot_base]) / a.chunk) + (size_t)a.mu * a.H + 16);
    std::vector<std::vector<HotUnit>> seg((size_t)a.rounds * a.IB);       // the (round, item block) segments of one sub-stripe
    std::vector<size_t> place;
    for (int sa = 0; sa < a.mu; sa++) {
        for (auto& v : seg) v.clear();
        for (int ib = 0; ib < a.IB; ib++)
            for (int hx = a.hot_block_lo[ib]; hx < a.hot_block_lo[ib + 1]; hx++) {
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
                    std::vector<HotUnit>& out = seg[(size_t)rnd * a.IB + ib];
                    for (int64_t pc = 0; pc < pieces; pc++) {
                        HotUnit u{};
                        u.bstart = a.block_off[blk];
                        u.bn = (int32_t)bn;
                        u.bid = (uint32_t)((size_t)a.member * a.n_blocks + blk);
                        u.start = lo + n * pc / pieces;
                        u.count = (int32_t)(lo + n * (pc + 1) / pieces - u.start);
                        u.item = a.hot_items[hx];
                        u.weight = 1.0f / (float)pieces;
                        out.push_back(u);
                    }
                }
            }
        for (int rnd = 0; rnd < a.rounds; rnd++)
            for (int ib = 0; ib < a.IB; ib++) {
                const std::vector<HotUnit>& v = seg[(size_t)rnd * a.IB + ib];
                visit_units[((size_t)sa * a.rounds + rnd) * a.IB + ib] = (int)units.size();
                const size_t base = units.size();
                units.resize(base + v.size());
                place.assign((size_t)a.chunk + 2, 0);                    // stable counting sort, key 0 = longest run
                for (const HotUnit& u : v) place[(size_t)(a.chunk - u.count) + 1]++;
                for (size_t c = 1; c < place.size(); c++) place[c] += place[c - 1];
                for (const HotUnit& u : v) units[base + place[(size_t)(a.chunk - u.count)]++] = u;
            }
    }
    visit_units.back() = (int)units.size();
}

}  // namespace mfsgd
