// tmap.cu -- see tmap.hpp
#include "tmap.hpp"

#include "fv2d.cuh"

namespace hrw {

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static encode_tiled_fn encode_tiled() {
   static encode_tiled_fn fn = [] {
      void *p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
      return (encode_tiled_fn)p;
   }();
   return fn;
}

int fv_tmap_2d(Fv *fv, const double *cell0, uint32_t box0, uint32_t box1, CUtensorMap *out) {
   for (const auto &e : fv->tmaps)
      if (e.ptr == cell0 && e.box0 == box0 && e.box1 == box1) {
         *out = e.map;
         return HRWENO_OK;
      }
   encode_tiled_fn enc = encode_tiled();
   if (!enc) return fail(HRWENO_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
   Fv::TmapEntry e;
   e.ptr = cell0;
   e.box0 = box0;
   e.box1 = box1;
   void *base = const_cast<double *>(cell0) - (int64_t)PAD2 * fv->pitch - PAD;
   const cuuint64_t dims[2] = {(cuuint64_t)fv->pitch, (cuuint64_t)fv->nrows_alloc};
   const cuuint64_t strides[1] = {(cuuint64_t)fv->pitch * sizeof(double)};
   const cuuint32_t box[2] = {box0, box1};
   const cuuint32_t estr[2] = {1, 1};
   const CUresult r = enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   if (r != CUDA_SUCCESS) return fail(HRWENO_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
   if (fv->tmaps.size() >= 64) fv->tmaps.erase(fv->tmaps.begin()); // states come and go with the integrators: bounded cache
   fv->tmaps.push_back(e);
   *out = e.map;
   return HRWENO_OK;
}

} // namespace hrw
