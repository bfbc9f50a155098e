// What the device headers need from the toolchain, for both compilers that build them:
//   nvcc  (libpb2.so, offline)            -- the standard headers;
//   NVRTC (user-defined targets, at run time, pb2_user.cu) -- no host headers are reachable there: fixed-width
//         integers and INFINITY are spelled out, the CUDA built-ins are predeclared by the compiler.
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#else
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#endif

// shared-memory plan functions of the targets: called by the host launcher, and (wrappers delegating to their base)
// parsed by the run-time compiler, which accepts no unannotated functions
#define PB2_HOSTFN __host__ __device__
