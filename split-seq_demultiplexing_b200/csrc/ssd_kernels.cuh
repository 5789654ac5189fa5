// ssd_kernels.cuh — the CUDA kernels of the --version fast hot path (sm_100a), one header per stage.
#pragma once
#include "ssd_device.cuh"
#include "ssd_sort.cuh"
#include "ssd_onesweep.cuh"
#include "ssd_scan.cuh"
#include "ssd_emit.cuh"
#include "ssd_emit_tiled.cuh"
