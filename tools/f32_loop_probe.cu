// Developer probe: compile ONLY the benchmark force kernels (fp32, FLOAT32 mode, D=3/2, uniform masses) so that
// cuobjdump can show their inner loop for a given -DNB_F32_ACC_VARIANT / -DNB_F32_UNROLL (tools/f32_loop_probe.sh).
#define NB_TUNE_HARNESS
#include "../nbody_cosmological_simulation_b200/csrc/accel.cu"
#ifndef NB_F32_UNROLL
#define NB_F32_UNROLL 4
#endif
template __global__ void nb::accel_kernel<nb::ForceF32<3, nb::Q_F32, 2, 256, true, NB_F32_UNROLL>, 0, false>(const nb::AccelArgs);
template __global__ void nb::accel_kernel<nb::ForceF32<2, nb::Q_F32, 2, 256, true, NB_F32_UNROLL>, 0, false>(const nb::AccelArgs);
template __global__ void nb::accel_kernel<nb::ForceF32<3, nb::Q_F32, 2, 256, false, NB_F32_UNROLL>, 0, false>(const nb::AccelArgs);
