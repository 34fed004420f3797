// factorize_demo -- smallest end-to-end use of the C++ host mirror: 2000 random ratings, 5 epochs on the GPU.
//   ./host/factorize_demo [path/to/libmfsgd.so]        exit 0 on success; exit 3 when no GPU is usable.
#include <cstdio>
#include <cstdlib>

#include "MatrixFactorizationSGD.hpp"

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "matrixfactorizationsgd.java_b200/lib/libmfsgd.so";
    try {
        MatrixFactorizationSGD mf(path);
        const int nU = 200, nI = 100, n = 2000, k = 32;
        std::vector<int32_t> u(n), i(n);
        std::vector<float> r(n);
        uint32_t s = 12345u;
        auto next = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
        for (int t = 0; t < n; t++) { u[t] = next() % nU; i[t] = next() % nI; r[t] = 1.0f + (next() % 5); }
        bool threw = false;
        try { mf.factorize(u, i, r, nU, nI, 0, 0.01f, 0.05f, 1, 1); } catch (const std::invalid_argument&) { threw = true; }
        if (!threw) { fprintf(stderr, "bad shape was not rejected\n"); return 1; }
        MatrixFactorizationSGD::Factors f0, f5;
        try {
            f0 = mf.factorize(u, i, r, nU, nI, k, 0.02f, 0.05f, 0, 7);
            f5 = mf.factorize(u, i, r, nU, nI, k, 0.02f, 0.05f, 5, 7);
        } catch (const std::runtime_error& e) {
            fprintf(stderr, "GPU path unavailable: %s\n", e.what());
            return 3;
        }
        const double before = mf.rmse(f0, u, i, r), after = mf.rmse(f5, u, i, r);
        printf("train RMSE %.4f -> %.4f after 5 epochs\n", before, after);
        return after < before ? 0 : 1;
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
