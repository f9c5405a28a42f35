// nbmf_glue.h -- what the reference's src/*.cpp need to route the hot path through libnbmf_cuda
// (INTEGRATION.md).  Header-only; include after <RcppArmadillo.h>.
#pragma once
#include "nbmf_cuda.h"

#include <cstdlib>
#include <string>
#include <vector>

// Devices of this R process: the environment variable NBMF_DEVICES ("0,1,2,3", or "all" for every visible
// GPU up to 8); unset = device 0 only.  More than one device makes VB_Aspect_Rcpp shard the columns of V over
// them inside the library (nbmf_group_*: one NCCL rank and one host thread per device).
inline std::vector<int> nbmf_glue_devices()
{
    std::vector<int> d;
    const char *e = std::getenv("NBMF_DEVICES");
    if (!e || !*e) { d.push_back(0); return d; }
    std::string s(e);
    if (s == "all") {
        int n = nbmf_device_count();
        for (int i = 0; i < n && i < 8; i++) d.push_back(i);
        if (d.empty()) d.push_back(0);
        return d;
    }
    size_t pos = 0;
    while (pos < s.size()) {
        size_t c = s.find(',', pos);
        if (c == std::string::npos) c = s.size();
        if (c > pos) d.push_back(std::atoi(s.substr(pos, c - pos).c_str()));
        pos = c + 1;
    }
    if (d.empty()) d.push_back(0);
    return d;
}

struct NbmfHandle {                       // RAII so that Rcpp::stop() cannot leak the handle
    nbmf_handle *h = nullptr;
    explicit NbmfHandle(int precision = NBMF_PREC_AUTO, int planned_iters = 0)
    {
        nbmf_opts o = {};
        o.device = nbmf_glue_devices()[0]; o.precision = precision; o.planned_iters = planned_iters;
        if (nbmf_create(&h, &o) != NBMF_OK) Rcpp::stop("libnbmf_cuda: no CUDA device (there is no CPU fallback)");
    }
    ~NbmfHandle() { nbmf_destroy(h); }
    NbmfHandle(const NbmfHandle &) = delete;
    NbmfHandle &operator=(const NbmfHandle &) = delete;
    void ck(int rc) { if (rc != NBMF_OK) Rcpp::stop(nbmf_last_error(h)); }   // status code -> R error (BEGIN_RCPP / END_RCPP)
};

struct NbmfGroup {                        // the same for a group of devices (column sharding inside the library)
    nbmf_group *g = nullptr;
    NbmfGroup(const std::vector<int> &devices, int precision, int planned_iters)
    {
        nbmf_opts o = {};
        o.precision = precision; o.planned_iters = planned_iters;
        if (nbmf_group_create(&g, &o, devices.data(), (int)devices.size()) != NBMF_OK)
            Rcpp::stop("libnbmf_cuda: could not open the devices of NBMF_DEVICES");
    }
    ~NbmfGroup() { nbmf_group_destroy(g); }
    NbmfGroup(const NbmfGroup &) = delete;
    NbmfGroup &operator=(const NbmfGroup &) = delete;
    void ck(int rc) { if (rc != NBMF_OK) Rcpp::stop(nbmf_group_last_error(g)); }
};
