// factorize_demo -- smallest end-to-end use of the C++ host mirror: 2000 random ratings, 5 epochs on the GPU, through every
// entry point the host mirrors (factorize, factorizeMixed, factorizeModel, factorizeEarlyStop, rmse, rmseModel).
//   ./host/factorize_demo [path/to/libmfsgd.so]        exit 0 on success; exit 3 when no GPU is usable.
#include <cmath>
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
        threw = false;       // stand-in :443: binary16 rows need k % 4 == 0
        try { mf.factorizeMixed(u, i, r, nU, nI, 30, 0.01f, 0.05f, 1, 1); } catch (const std::invalid_argument&) { threw = true; }
        if (!threw) { fprintf(stderr, "factorizeMixed: bad shape was not rejected\n"); return 1; }
        threw = false;       // stand-in :356: lrDecay in (0, 1], patience >= 0, minDelta in [0, 1)
        try { mf.factorizeEarlyStop(u, i, r, u, i, r, nU, nI, k, 0.01f, 0.05f, 3, 1, true, true, 1.5f, 1, 0.0f); } catch (const std::invalid_argument&) { threw = true; }
        if (!threw) { fprintf(stderr, "factorizeEarlyStop: bad schedule was not rejected\n"); return 1; }
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
        if (!(after < before)) return 1;
        // binary16 rows of P (stand-in factorizeMixed): trains as well, P comes back as binary32
        const MatrixFactorizationSGD::Factors h5 = mf.factorizeMixed(u, i, r, nU, nI, k, 0.02f, 0.05f, 5, 7);
        const double mixed = mf.rmse(h5, u, i, r);
        printf("binary16 rows of P: train RMSE %.4f after 5 epochs\n", mixed);
        if (!(mixed < before) || std::fabs(mixed - after) > 0.05 * after) return 1;
        // model extension (stand-in factorizeModel / rmseModel): global mean + biases
        const MatrixFactorizationSGD::Model m0 = mf.factorizeModel(u, i, r, nU, nI, k, 0.02f, 0.05f, 0, 7, true, true);
        const MatrixFactorizationSGD::Model m5 = mf.factorizeModel(u, i, r, nU, nI, k, 0.02f, 0.05f, 5, 7, true, true);
        const double mb = mf.rmseModel(m0, u, i, r), ma = mf.rmseModel(m5, u, i, r);
        printf("extended model: global mean %.4f, train RMSE %.4f -> %.4f after 5 epochs\n", m5.globalMean, mb, ma);
        if (!(ma < mb) || !(m5.globalMean > 2.5f && m5.globalMean < 3.5f) || m5.userBias.size() != (size_t)nU) return 1;
        // schedule + early stopping (stand-in factorizeEarlyStop): the last 200 triplets validate
        const std::vector<int32_t> tu(u.begin(), u.end() - 200), ti(i.begin(), i.end() - 200), vu(u.end() - 200, u.end()), vi(i.end() - 200, i.end());
        const std::vector<float> tr(r.begin(), r.end() - 200), vr(r.end() - 200, r.end());
        const MatrixFactorizationSGD::EarlyStopResult es =
            mf.factorizeEarlyStop(tu, ti, tr, vu, vi, vr, nU, nI, k, 0.05f, 0.05f, 12, 7, true, true, 0.9f, 2, 0.001f);
        printf("early stopping: %d of 12 epochs run, validation RMSE %.4f -> %.4f\n", es.epochsRun,
               es.validationRmse.empty() ? 0.0 : es.validationRmse.front(), es.validationRmse.empty() ? 0.0 : es.validationRmse.back());
        if (es.epochsRun < 1 || es.epochsRun > 12 || (int)es.validationRmse.size() != es.epochsRun) return 1;
        for (double v : es.validationRmse)
            if (!(v > 0.0) || !std::isfinite(v)) return 1;
        if (std::fabs(mf.rmseModel(es.model, vu, vi, vr) - es.validationRmse.back()) > 1e-4 * es.validationRmse.back()) return 1;
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
