// MatrixFactorizationSGD.hpp -- C++ host mirror of the reference's entry point over the C ABI.
//
// The production host is Java (java/MatrixFactorizationSGDGpu.java, Panama FFM); this image has no JDK, so
// the same thin layer exists in C++ (compiled and run by `make host` / tests) and in Python
// (matrixfactorizationsgd.java_b200/host.py). Same name, argument order and error behaviour as the stand-in
// baseline/java/MatrixFactorizationSGD.java:109 (factorize) and :169 (rmse): bad shapes throw
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

    void* lib_ = nullptr;
    int (*config_default_)(mfsgd_config*) = nullptr;
    const char* (*last_error_)() = nullptr;
    int (*factorize_)(const int32_t*, const int32_t*, const float*, int64_t, const mfsgd_config*, int32_t, float*, float*) = nullptr;
    int (*create_)(const mfsgd_config*, mfsgd_handle**) = nullptr;
    void (*destroy_)(mfsgd_handle*) = nullptr;
    int (*load_ratings_)(mfsgd_handle*, const int32_t*, const int32_t*, const float*, int64_t) = nullptr;
    int (*set_factors_)(mfsgd_handle*, const float*, const float*) = nullptr;
    int (*rmse_)(mfsgd_handle*, const int32_t*, const int32_t*, const float*, int64_t, double*) = nullptr;
    int (*read_ratings_)(const char*, int32_t, mfsgd_ratings*) = nullptr;
    void (*free_ratings_)(mfsgd_ratings*) = nullptr;
};
