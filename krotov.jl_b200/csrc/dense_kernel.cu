// placeholder until the DMMA engine lands: the dense path reports "unsupported" loudly
#include "dense_kernel.cuh"
namespace kr {
struct DenseEngine { int dummy; };
DenseEngine *dense_create(int, int, int, int, int, const std::vector<std::complex<double>> &, const std::vector<int> &,
                          const double *, const double *, int, cudaStream_t, std::string &err) {
    err = "dense path (d > 32) not built yet";
    return nullptr;
}
void dense_destroy(DenseEngine *e) { delete e; }
void dense_info(DenseEngine *, krotov_info *) {}
bool dense_set_cheby(DenseEngine *, int, int, const std::vector<int> &, const std::vector<double> &,
                     const std::vector<double> &, const std::vector<int> &, const std::vector<double> &, int,
                     const std::vector<std::complex<double>> &, std::string &) { return false; }
bool dense_forward(DenseEngine *, const double *, double2 *, long long &, std::string &) { return false; }
bool dense_iterate(DenseEngine *, const double *, double *, const double *, const double *, double *, const double2 *,
                   double2 *, long long &, std::string &) { return false; }
bool dense_set_chi(DenseEngine *, const double *, std::string &) { return false; }
bool dense_get_states(DenseEngine *, double *, std::string &) { return false; }
bool dense_get_storage(DenseEngine *, int, int, int, int, double *, std::string &) { return false; }
}  // namespace kr
