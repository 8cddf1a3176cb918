// One explicit (dimension, smoothing kernel) instantiation of the engine per
// translation unit: compiled as  nvcc -DTIT_D=<2|3> -DTIT_K=<0..5> inst.cu
#include "engine.cuh"

#if !defined(TIT_D) || !defined(TIT_K)
#error "compile with -DTIT_D=<dim> -DTIT_K=<kernel id>"
#endif

namespace titgpu {
namespace {
struct Registrar {
  Registrar() { register_engine(TIT_D, TIT_K, Engine<TIT_D, TIT_K>::vtable()); }
} registrar_;
}  // namespace
}  // namespace titgpu
