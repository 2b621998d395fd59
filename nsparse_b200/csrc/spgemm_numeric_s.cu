// fp32 instantiation of the numeric SpGEMM phase
#include "spgemm_numeric.cuh"
namespace nsp {
template int spgemm_numeric<float>(nsp_context *, int, int, int, const int *, const int *, const float *,
                                   const int *, const int *, const float *, const long long *, int *,
                                   float *, int, int);
template int spgemm_numeric_reserve<float>(nsp_context *, int, long long, long long, int, int);
}
