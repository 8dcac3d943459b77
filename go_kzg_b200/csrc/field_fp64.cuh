// field_fp64.cuh -- FP64-pipe Montgomery product; the code lives in field_fp64_impl.cuh, which
// field.cuh includes once the Fp type exists.
#pragma once
#include "field.cuh"
