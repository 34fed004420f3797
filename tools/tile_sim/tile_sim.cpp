// tile_sim.cpp -- design study for DESIGN.md section 8.1 (NOT product code, NOT the oracle): a CPU event simulation of the
// "user tile" kernel's SGD semantics, to learn whether it converges like the sequential reference before anyone writes it.
//
// Model. Users are cut into tiles of U rows; a CTA owns one tile at a time (its P rows are private to it while it works,
// so P is always current) and its warps walk the tile's records, which are sorted by item, in chunks of CH records; inside
// a chunk a maximal group of records of one item is a RUN: the warp snapshots q_i from the global Q when the run starts,
// applies the run's ratings one after another (the reference update rule, fp32) with q_i private, and when the run ENDS
// -- `len + overhead` time units later -- merges w_i * (q_i_final - q_i_snapshot) into the global Q. Runs of one item that
// overlap in time therefore start from the same stale q_i; `weights` is the merge rule under test.
// `ctas * warps` warps run concurrently; every warp's clock advances by the length of the runs it walks.
//
// build: g++ -O2 -ffp-contract=off -shared -fPIC -o libtilesim.so tile_sim.cpp      (driver: tools/tile_sim/run.py)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

namespace {

struct Warp {
    int cta = 0;
    // the run in flight (merge pending at its end)
    int item = -1;
    std::vector<float> q, q0;
    // the chunk being walked
    int64_t pos = 0, end = 0;
};
struct Event {
    double t;
    int warp;
    bool operator<(const Event& o) const { return t > o.t; }   // min-heap
};

int g_gpu_arith = 0;

}  // namespace

extern "C" void tilesim_set_gpu_arithmetic(int on) { g_gpu_arith = on; }

// records sorted by (tile, item order); tile_off[n_tiles+1]; tile_order[n_tiles] = this epoch's visiting order
extern "C" int tilesim_epoch(const int32_t* u, const int32_t* it, const float* r, const int64_t* tile_off, int32_t n_tiles,
                             const int32_t* tile_order, float* P, float* Q, int32_t k, float lr, float lambda, int32_t ctas,
                             int32_t warps_per_cta, int32_t chunk, double run_overhead, const float* item_weight,
                             int32_t store_instead_of_add, double* stats /* [0] runs, [1] mean concurrent runs per merge on the same item */) {
    const int W = ctas * warps_per_cta;
    std::vector<Warp> warps((size_t)W);
    for (int w = 0; w < W; w++) {
        warps[(size_t)w].cta = w / warps_per_cta;
        warps[(size_t)w].q.resize((size_t)k);
        warps[(size_t)w].q0.resize((size_t)k);
    }
    // per CTA: the tile it holds and the next unclaimed record of that tile
    std::vector<int64_t> cta_next((size_t)ctas, 0), cta_end((size_t)ctas, 0);
    int next_tile = 0;
    auto cta_claim_tile = [&](int c) -> bool {
        while (next_tile < n_tiles) {
            const int t = tile_order[next_tile++];
            if (tile_off[t + 1] > tile_off[t]) {
                cta_next[(size_t)c] = tile_off[t];
                cta_end[(size_t)c] = tile_off[t + 1];
                return true;
            }
        }
        return false;
    };
    auto warp_claim_chunk = [&](Warp& w) -> bool {
        const int c = w.cta;
        if (cta_next[(size_t)c] >= cta_end[(size_t)c] && !cta_claim_tile(c)) return false;
        w.pos = cta_next[(size_t)c];
        w.end = std::min(cta_end[(size_t)c], w.pos + chunk);
        cta_next[(size_t)c] = w.end;
        return true;
    };
    std::vector<int> in_flight;   // runs in flight per item (statistics)
    int max_item = 0;
    for (int64_t j = 0; j < tile_off[n_tiles]; j++) max_item = std::max(max_item, it[j]);
    in_flight.assign((size_t)max_item + 1, 0);
    double n_runs = 0, conc_sum = 0;
    std::priority_queue<Event> heap;
    // walks the warp's next run now (time t): snapshot, apply, and report when it ends; false if the warp has no work left
    auto start_run = [&](int wi, double t) -> bool {
        Warp& w = warps[(size_t)wi];
        if (w.pos >= w.end && !warp_claim_chunk(w)) return false;
        const int item = it[w.pos];
        w.item = item;
        float* qg = Q + (int64_t)item * k;
        std::memcpy(w.q0.data(), qg, sizeof(float) * k);
        std::memcpy(w.q.data(), qg, sizeof(float) * k);
        int64_t len = 0;
        while (w.pos < w.end && it[w.pos] == item) {
            float* p = P + (int64_t)u[w.pos] * k;
            float* q = w.q.data();
            if (g_gpu_arith) {
                // the GPU kernels' FAST arrangement at k = 128, one float4 chunk per lane, 32 lanes (update_math.cuh;
                // oracle.cpp ORC_ORDER_WARP_TREE_FMA): makes the simulation with one CTA of one warp the bit-exact twin
                // of tools/tile_kernel_draft/tile_kernel.cu run in its checking mode
                float s[32];
                for (int l = 0; l < 32; l++) {
                    float lo = 0.f, hi = 0.f;
                    bool first = true;
                    for (int c = l; c < k / 4; c += 32) {
                        const float* pp = p + 4 * c;
                        const float* qq = q + 4 * c;
                        if (first) { lo = pp[0] * qq[0]; hi = pp[1] * qq[1]; first = false; }
                        else { lo = std::fmaf(pp[0], qq[0], lo); hi = std::fmaf(pp[1], qq[1], hi); }
                        lo = std::fmaf(pp[2], qq[2], lo);
                        hi = std::fmaf(pp[3], qq[3], hi);
                    }
                    s[l] = lo + hi;
                }
                for (int m = 16; m >= 1; m >>= 1) {
                    float t2[32];
                    for (int l = 0; l < 32; l++) t2[l] = s[l] + s[l ^ m];
                    for (int l = 0; l < 32; l++) s[l] = t2[l];
                }
                const float e = r[w.pos] - s[0];
                const float a_ = 1.0f - lr * lambda, b_ = lr * e;
                for (int f = 0; f < k; f++) {
                    const float pf = p[f], qf = q[f];
                    p[f] = std::fmaf(b_, qf, a_ * pf);
                    q[f] = std::fmaf(b_, pf, a_ * qf);
                }
            } else {
                float dot = 0.0f;
                for (int f = 0; f < k; f++) dot = dot + p[f] * q[f];
                const float e = r[w.pos] - dot;
                for (int f = 0; f < k; f++) {
                    const float pf = p[f], qf = q[f];
                    p[f] = pf + lr * (e * qf - lambda * pf);
                    q[f] = qf + lr * (e * pf - lambda * qf);
                }
            }
            w.pos++;
            len++;
        }
        in_flight[(size_t)item]++;
        heap.push({t + (double)len + run_overhead, wi});
        return true;
    };
    for (int wi = 0; wi < W; wi++) start_run(wi, 0.0);
    while (!heap.empty()) {
        const Event ev = heap.top();
        heap.pop();
        Warp& w = warps[(size_t)ev.warp];
        // merge the finished run
        float* qg = Q + (int64_t)w.item * k;
        // store_instead_of_add == 2: dynamic rule -- average over the runs that are in flight on the item right now
        // (what a per-item atomic counter would tell the kernel), instead of the static expectation in item_weight
        const float wt = store_instead_of_add == 2 ? 1.0f / (float)in_flight[(size_t)w.item] : (item_weight ? item_weight[w.item] : 1.0f);
        if (store_instead_of_add == 1) {
            std::memcpy(qg, w.q.data(), sizeof(float) * k);
        } else {
            for (int f = 0; f < k; f++) qg[f] = qg[f] + wt * (w.q[(size_t)f] - w.q0[(size_t)f]);
        }
        n_runs += 1;
        conc_sum += in_flight[(size_t)w.item];
        in_flight[(size_t)w.item]--;
        start_run(ev.warp, ev.t);
    }
    if (stats) {
        stats[0] = n_runs;
        stats[1] = n_runs > 0 ? conc_sum / n_runs : 0.0;
    }
    return 0;
}
