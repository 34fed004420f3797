// MatrixFactorizationSGD.hpp -- C++ host mirror of the reference's entry point over the C ABI.
//
// The production host is Java (java/MatrixFactorizationSGDGpu.java, Panama FFM); this image has no JDK, so
// the same thin layer exists in C++ (compiled and run by `make host` / tests) and in Python
// (matrixfactorizationsgd.java_b200/host.py). Same name, argument order and error behaviour as the stand-in
// baseline/java/MatrixFactorizationSGD.java:109 (factorize), :169 (rmse), :305 (factorizeModel), :350 (factorizeEarlyStop),
// :389 (rmseModel) and :439 (factorizeMixed): bad shapes and schedules throw
// std::invalid_argument before any GPU work; a failing library call throws std::runtime_error with
// mfsgd_last_error(). Binds libmfsgd.so with dlopen/dlsym only -- what FFM's downcall handles do.
#pragma once
#include <dlfcn.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/mfsgd.h"

class MatrixFactorizationSGD {
public:
    struct Factors {
        std::vector<float> P, Q;   // row-major nUsers x k, nItems x k
        int nUsers, nItems, k;
    };

    explicit MatrixFactorizationSGD(const std::string& lib_path = "libmfsgd.so") {
        lib_ = dlopen(lib_path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!lib_) throw std::runtime_error(std::string("cannot load libmfsgd.so: ") + dlerror());
        bind(config_default_, "mfsgd_config_default");
        bind(last_error_, "mfsgd_last_error");
        bind(factorize_, "mfsgd_factorize");
        bind(create_, "mfsgd_create");
        bind(destroy_, "mfsgd_destroy");
        bind(load_ratings_, "mfsgd_load_ratings");
        bind(set_factors_, "mfsgd_set_factors");
        bind(rmse_, "mfsgd_rmse");
        bind(load_heldout_, "mfsgd_load_heldout");
        bind(init_factors_, "mfsgd_init_factors");
        bind(train_, "mfsgd_train");
        bind(get_factors_, "mfsgd_get_factors");
        bind(get_model_, "mfsgd_get_model");
        bind(set_biases_, "mfsgd_set_biases");
        bind(get_progress_, "mfsgd_get_progress");
        bind(set_eval_every_epoch_, "mfsgd_set_eval_every_epoch");
        bind(read_ratings_, "mfsgd_read_ratings");
        bind(free_ratings_, "mfsgd_free_ratings");
    }
    ~MatrixFactorizationSGD() {
        if (lib_) dlclose(lib_);
    }
    MatrixFactorizationSGD(const MatrixFactorizationSGD&) = delete;
    MatrixFactorizationSGD& operator=(const MatrixFactorizationSGD&) = delete;

    // stand-in line 109; mode/n_gpus select the GPU execution mode (MFSGD_MODE_*)
    Factors factorize(const std::vector<int32_t>& users, const std::vector<int32_t>& items, const std::vector<float>& ratings,
                      int nUsers, int nItems, int k, float lr, float lambda, int epochs, uint64_t seed,
                      int mode = MFSGD_MODE_HOGWILD, int n_gpus = 1) const {
        if (users.size() != items.size() || users.size() != ratings.size())
            throw std::invalid_argument("triplet arrays differ in length");
        if (k <= 0 || nUsers <= 0 || nItems <= 0 || epochs < 0) throw std::invalid_argument("bad shape");
        mfsgd_config cfg = config(nUsers, nItems, k, lr, lambda, seed, mode, n_gpus);
        Factors f{std::vector<float>((size_t)nUsers * k), std::vector<float>((size_t)nItems * k), nUsers, nItems, k};
        check(factorize_(users.data(), items.data(), ratings.data(), (int64_t)ratings.size(), &cfg, epochs, f.P.data(), f.Q.data()));
        return f;
    }

    // stand-in line 439: the rows of P kept as binary16 on the device (stochastic rounding from a counter hash), arithmetic in
    // binary32; P comes back widened exactly
    Factors factorizeMixed(const std::vector<int32_t>& users, const std::vector<int32_t>& items, const std::vector<float>& ratings,
                           int nUsers, int nItems, int k, float lr, float lambda, int epochs, uint64_t seed) const {
        if (users.size() != items.size() || users.size() != ratings.size())
            throw std::invalid_argument("triplet arrays differ in length");
        if (k <= 0 || k % 4 != 0 || nUsers <= 0 || nItems <= 0 || epochs < 0) throw std::invalid_argument("bad shape");
        mfsgd_config cfg = config(nUsers, nItems, k, lr, lambda, seed, MFSGD_MODE_HOGWILD, 1);
        cfg.p_storage = MFSGD_STORAGE_F16;
        Factors f{std::vector<float>((size_t)nUsers * k), std::vector<float>((size_t)nItems * k), nUsers, nItems, k};
        check(factorize_(users.data(), items.data(), ratings.data(), (int64_t)ratings.size(), &cfg, epochs, f.P.data(), f.Q.data()));
        return f;
    }

    // stand-in lines 257 / 338: factors plus the extension's terms; userBias / itemBias are empty when the biases are off
    struct Model {
        std::vector<float> P, Q, userBias, itemBias;
        float globalMean = 0.f;
        int nUsers = 0, nItems = 0, k = 0;
        bool biased = false;
    };
    struct EarlyStopResult {
        Model model;
        int epochsRun = 0;
        std::vector<double> validationRmse;
    };

    // stand-in line 305: r ~ mu + b_u + b_i + p_u . q_i
    Model factorizeModel(const std::vector<int32_t>& users, const std::vector<int32_t>& items, const std::vector<float>& ratings,
                         int nUsers, int nItems, int k, float lr, float lambda, int epochs, uint64_t seed, bool useGlobalMean,
                         bool useBiases) const {
        return trainModel(users, items, ratings, nullptr, nullptr, nullptr, nUsers, nItems, k, lr, lambda, epochs, seed, useGlobalMean,
                          useBiases, 1.0f, 0, 0.0f)
            .model;
    }

    // stand-in line 350: learning-rate schedule lr_(e+1) = lr_e * lrDecay and early stopping on the validation RMSE
    EarlyStopResult factorizeEarlyStop(const std::vector<int32_t>& users, const std::vector<int32_t>& items, const std::vector<float>& ratings,
                                       const std::vector<int32_t>& vUsers, const std::vector<int32_t>& vItems,
                                       const std::vector<float>& vRatings, int nUsers, int nItems, int k, float lr, float lambda,
                                       int maxEpochs, uint64_t seed, bool useGlobalMean, bool useBiases, float lrDecay, int patience,
                                       float minDelta) const {
        if (!(lrDecay > 0.0f) || lrDecay > 1.0f || patience < 0 || !(minDelta >= 0.0f) || minDelta >= 1.0f)
            throw std::invalid_argument("bad schedule");
        if (vUsers.size() != vItems.size() || vUsers.size() != vRatings.size())
            throw std::invalid_argument("triplet arrays differ in length");
        return trainModel(users, items, ratings, &vUsers, &vItems, &vRatings, nUsers, nItems, k, lr, lambda, maxEpochs, seed, useGlobalMean,
                          useBiases, lrDecay, patience, minDelta);
    }

    // stand-in line 389: e = (r - mu) - ((p_u . q_i + b_u) + b_i), evaluated by the RMSE kernel
    double rmseModel(const Model& m, const std::vector<int32_t>& users, const std::vector<int32_t>& items,
                     const std::vector<float>& ratings) const {
        std::vector<float> centred(ratings.size());
        for (size_t t = 0; t < ratings.size(); t++) centred[t] = ratings[t] - m.globalMean;     // one binary32 subtraction, as :399
        mfsgd_config cfg = config(m.nUsers, m.nItems, m.k, 1e-3f, 0.f, 0, MFSGD_MODE_HOGWILD, 1);
        cfg.model = m.biased ? MFSGD_MODEL_BIASES : 0u;
        mfsgd_handle* h = nullptr;
        check(create_(&cfg, &h));
        double out = 0.0;
        int rc = load_ratings_(h, nullptr, nullptr, nullptr, 0);
        if (rc == MFSGD_OK) rc = set_factors_(h, m.P.data(), m.Q.data());
        if (rc == MFSGD_OK && m.biased) rc = set_biases_(h, m.userBias.data(), m.itemBias.data());
        if (rc == MFSGD_OK) rc = rmse_(h, users.data(), items.data(), centred.data(), (int64_t)centred.size(), &out);
        const std::string msg = rc == MFSGD_OK ? std::string() : std::string(last_error_());
        destroy_(h);
        if (rc != MFSGD_OK) throw std::runtime_error("mfsgd error " + std::to_string(rc) + ": " + msg);
        return out;
    }

    // A ratings file (MovieLens u.data / ratings.csv / ratings.dat, Netflix-Prize text) as the triplets factorize takes;
    // userIds[u] / itemIds[i] give the file's id of dense row u / i. Host-only (mfsgd_read_ratings).
    struct RatingsFile {
        std::vector<int32_t> users, items;
        std::vector<float> ratings;
        std::vector<int64_t> userIds, itemIds;
        int nUsers = 0, nItems = 0;
    };
    RatingsFile readRatings(const std::string& path, int format = MFSGD_FORMAT_AUTO) const {
        mfsgd_ratings r;
        check(read_ratings_(path.c_str(), format, &r));
        RatingsFile f;
        f.users.assign(r.users, r.users + r.n);
        f.items.assign(r.items, r.items + r.n);
        f.ratings.assign(r.ratings, r.ratings + r.n);
        f.userIds.assign(r.user_ids, r.user_ids + r.n_users);
        f.itemIds.assign(r.item_ids, r.item_ids + r.n_items);
        f.nUsers = r.n_users;
        f.nItems = r.n_items;
        free_ratings_(&r);
        return f;
    }

    // stand-in line 169
    double rmse(const Factors& f, const std::vector<int32_t>& users, const std::vector<int32_t>& items,
                const std::vector<float>& ratings) const {
        mfsgd_config cfg = config(f.nUsers, f.nItems, f.k, 1e-3f, 0.f, 0, MFSGD_MODE_HOGWILD, 1);
        mfsgd_handle* h = nullptr;
        check(create_(&cfg, &h));
        double out = 0.0;
        int rc = load_ratings_(h, nullptr, nullptr, nullptr, 0);
        if (rc == MFSGD_OK) rc = set_factors_(h, f.P.data(), f.Q.data());
        if (rc == MFSGD_OK) rc = rmse_(h, users.data(), items.data(), ratings.data(), (int64_t)ratings.size(), &out);
        destroy_(h);
        check(rc);
        return out;
    }

private:
    template <typename F>
    void bind(F& fn, const char* name) {
        fn = reinterpret_cast<F>(dlsym(lib_, name));
        if (!fn) throw std::runtime_error(std::string("libmfsgd.so lacks ") + name);
    }
    void check(int rc) const {
        if (rc != MFSGD_OK) throw std::runtime_error("mfsgd error " + std::to_string(rc) + ": " + last_error_());
    }
    mfsgd_config config(int nUsers, int nItems, int k, float lr, float lambda, uint64_t seed, int mode, int n_gpus) const {
        mfsgd_config cfg;
        check(config_default_(&cfg));
        cfg.n_users = nUsers; cfg.n_items = nItems; cfg.k = k; cfg.lr = lr; cfg.lambda = lambda;
        cfg.seed = seed; cfg.mode = mode; cfg.n_gpus = n_gpus;
        return cfg;
    }

    EarlyStopResult trainModel(const std::vector<int32_t>& users, const std::vector<int32_t>& items, const std::vector<float>& ratings,
                               const std::vector<int32_t>* vUsers, const std::vector<int32_t>* vItems, const std::vector<float>* vRatings,
                               int nUsers, int nItems, int k, float lr, float lambda, int epochs, uint64_t seed, bool useGlobalMean,
                               bool useBiases, float lrDecay, int patience, float minDelta) const {
        if (users.size() != items.size() || users.size() != ratings.size())
            throw std::invalid_argument("triplet arrays differ in length");
        if (k <= 0 || nUsers <= 0 || nItems <= 0 || epochs < 0) throw std::invalid_argument("bad shape");
        mfsgd_config cfg = config(nUsers, nItems, k, lr, lambda, seed, MFSGD_MODE_HOGWILD, 1);
        cfg.model = (useGlobalMean ? MFSGD_MODEL_GLOBAL_MEAN : 0u) | (useBiases ? MFSGD_MODEL_BIASES : 0u);
        cfg.lr_decay = lrDecay;
        cfg.early_stop_patience = patience;
        cfg.early_stop_min_delta = minDelta;
        mfsgd_handle* h = nullptr;
        check(create_(&cfg, &h));
        EarlyStopResult res;
        Model& m = res.model;
        m.nUsers = nUsers; m.nItems = nItems; m.k = k; m.biased = useBiases;
        m.P.resize((size_t)nUsers * k);
        m.Q.resize((size_t)nItems * k);
        if (useBiases) {
            m.userBias.resize((size_t)nUsers);
            m.itemBias.resize((size_t)nItems);
        }
        std::vector<mfsgd_epoch_stats> stats((size_t)(vRatings ? epochs : 0));
        int32_t ran = 0;
        int rc = load_ratings_(h, users.data(), items.data(), ratings.data(), (int64_t)ratings.size());
        if (rc == MFSGD_OK && vRatings) rc = load_heldout_(h, vUsers->data(), vItems->data(), vRatings->data(), (int64_t)vRatings->size());
        if (rc == MFSGD_OK && vRatings) rc = set_eval_every_epoch_(h, 1);
        if (rc == MFSGD_OK) rc = init_factors_(h);
        if (rc == MFSGD_OK && epochs > 0) rc = train_(h, epochs, stats.empty() ? nullptr : stats.data());
        if (rc == MFSGD_OK) rc = get_progress_(h, &ran, nullptr, nullptr);
        if (rc == MFSGD_OK) rc = get_factors_(h, m.P.data(), m.Q.data());
        if (rc == MFSGD_OK) rc = get_model_(h, &m.globalMean, useBiases ? m.userBias.data() : nullptr, useBiases ? m.itemBias.data() : nullptr);
        const std::string msg = rc == MFSGD_OK ? std::string() : std::string(last_error_());
        destroy_(h);
        if (rc != MFSGD_OK) throw std::runtime_error("mfsgd error " + std::to_string(rc) + ": " + msg);
        res.epochsRun = ran;
        for (int e = 0; e < ran && e < (int)stats.size(); e++) res.validationRmse.push_back(stats[(size_t)e].heldout_rmse);
        return res;
    }

    void* lib_ = nullptr;
    int (*config_default_)(mfsgd_config*) = nullptr;
    const char* (*last_error_)() = nullptr;
    int (*factorize_)(const int32_t*, const int32_t*, const float*, int64_t, const mfsgd_config*, int32_t, float*, float*) = nullptr;
    int (*create_)(const mfsgd_config*, mfsgd_handle**) = nullptr;
    void (*destroy_)(mfsgd_handle*) = nullptr;
    int (*load_ratings_)(mfsgd_handle*, const int32_t*, const int32_t*, const float*, int64_t) = nullptr;
    int (*set_factors_)(mfsgd_handle*, const float*, const float*) = nullptr;
    int (*rmse_)(mfsgd_handle*, const int32_t*, const int32_t*, const float*, int64_t, double*) = nullptr;
    int (*load_heldout_)(mfsgd_handle*, const int32_t*, const int32_t*, const float*, int64_t) = nullptr;
    int (*init_factors_)(mfsgd_handle*) = nullptr;
    int (*train_)(mfsgd_handle*, int32_t, mfsgd_epoch_stats*) = nullptr;
    int (*get_factors_)(mfsgd_handle*, float*, float*) = nullptr;
    int (*get_model_)(mfsgd_handle*, float*, float*, float*) = nullptr;
    int (*set_biases_)(mfsgd_handle*, const float*, const float*) = nullptr;
    int (*get_progress_)(mfsgd_handle*, int32_t*, float*, int32_t*) = nullptr;
    int (*set_eval_every_epoch_)(mfsgd_handle*, int32_t) = nullptr;
    int (*read_ratings_)(const char*, int32_t, mfsgd_ratings*) = nullptr;
    void (*free_ratings_)(mfsgd_ratings*) = nullptr;
};
