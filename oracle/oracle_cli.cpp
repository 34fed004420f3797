/*
 * oracle_cli -- command-line front end of the CPU oracle (test infrastructure, see oracle.cpp).
 * Stands in for `java MatrixFactorizationSGD` (baseline/java/MatrixFactorizationSGD.java:241 main),
 * which cannot run in this image (no JDK).
 *
 *   oracle_cli [--users U --items I --ratings N --k K --epochs E --lr LR --lambda L]
 *              [--mode seq|threads=T] [--seed S]
 * Defaults are the ML-100K-shaped config. Prints one JSON line.
 */
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
float orc_default_init_scale(int k);
void orc_init_factors(float*, int64_t, int, uint64_t, uint64_t, float);
int orc_train(const int32_t*, const int32_t*, const float*, int64_t, float*, float*, int, int, int, float, float,
              int, int, uint64_t, int, int, float*);
int orc_train_hogwild(const int32_t*, const int32_t*, const float*, int64_t, float*, float*, int, int, int, float,
                      float, int, int, uint64_t, int, int, double*);
double orc_rmse(const float*, const float*, int, const int32_t*, const int32_t*, const float*, int64_t, int);
void orc_generate(uint64_t, int64_t, int64_t, int, int, int, double, int, double, int32_t*, int32_t*, float*,
                  uint8_t*, int);
int orc_hardware_threads(void);
}

int main(int argc, char** argv) {
    int nU = 943, nI = 1682, k = 32, epochs = 20, threads = 0;
    int64_t total = 100000;
    float lr = 0.01f, lambda = 0.05f;
    uint64_t seed = 20261018ULL;
    for (int a = 1; a < argc; a++) {
        std::string s = argv[a];
        auto next = [&]() -> const char* { return a + 1 < argc ? argv[++a] : "0"; };
        if (s == "--users") nU = atoi(next());
        else if (s == "--items") nI = atoi(next());
        else if (s == "--ratings") total = atoll(next());
        else if (s == "--k") k = atoi(next());
        else if (s == "--epochs") epochs = atoi(next());
        else if (s == "--lr") lr = (float)atof(next());
        else if (s == "--lambda") lambda = (float)atof(next());
        else if (s == "--seed") seed = strtoull(next(), nullptr, 10);
        else if (s == "--mode") {
            std::string m = next();
            if (m.rfind("threads=", 0) == 0) threads = atoi(m.c_str() + 8);
            else threads = 0;
        } else { fprintf(stderr, "unknown flag %s\n", s.c_str()); return 2; }
    }
    std::vector<int32_t> u(total), i(total);
    std::vector<float> r(total);
    std::vector<uint8_t> held(total);
    orc_generate(seed, 0, total, nU, nI, 2, 0.25, 3, 0.375, u.data(), i.data(), r.data(), held.data(),
                 orc_hardware_threads());
    std::vector<int32_t> tu, ti, hu, hi;
    std::vector<float> tr, hr;
    for (int64_t t = 0; t < total; t++) {
        if (held[t]) { hu.push_back(u[t]); hi.push_back(i[t]); hr.push_back(r[t]); }
        else { tu.push_back(u[t]); ti.push_back(i[t]); tr.push_back(r[t]); }
    }
    std::vector<float> P((size_t)nU * k), Q((size_t)nI * k);
    float scale = orc_default_init_scale(k);
    orc_init_factors(P.data(), nU, k, seed, 0, scale);
    orc_init_factors(Q.data(), nI, k, seed, 1, scale);
    int64_t n = (int64_t)tr.size();
    auto t0 = std::chrono::steady_clock::now();
    int rc;
    if (threads > 0)
        rc = orc_train_hogwild(tu.data(), ti.data(), tr.data(), n, P.data(), Q.data(), nU, nI, k, lr, lambda, 0,
                               epochs, seed, threads, 1, nullptr);
    else
        rc = orc_train(tu.data(), ti.data(), tr.data(), n, P.data(), Q.data(), nU, nI, k, lr, lambda, 0, epochs,
                       seed, 0, 1, nullptr);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc) { fprintf(stderr, "oracle failed\n"); return 1; }
    double heldout = orc_rmse(P.data(), Q.data(), k, hu.data(), hi.data(), hr.data(), (int64_t)hr.size(), 0);
    double train = orc_rmse(P.data(), Q.data(), k, tu.data(), ti.data(), tr.data(), n, 0);
    printf("{\"mode\": \"%s\", \"threads\": %d, \"host_cores\": %d, \"train_records\": %lld, \"heldout_records\": %lld, "
           "\"k\": %d, \"epochs\": %d, \"seconds\": %.6f, \"updates_per_sec\": %.6e, \"train_rmse\": %.6f, "
           "\"heldout_rmse\": %.6f}\n",
           threads > 0 ? "threads" : "seq", threads > 0 ? threads : 1, orc_hardware_threads(), (long long)n,
           (long long)hr.size(), k, epochs, secs, (double)n * epochs / secs, train, heldout);
    return 0;
}
