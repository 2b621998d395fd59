// fp64 instantiation of the numeric SpGEMM phase
#include "spgemm_numeric.cuh"
namespace nsp {
template int spgemm_numeric<double>(nsp_context *, int, int, int, const int *, const int *, const double *,
                                    const int *, const int *, const double *, const long long *, int *,
                                    double *, int, int);
template int spgemm_numeric_reserve<double>(nsp_context *, int, long long, long long, int, int);
}
