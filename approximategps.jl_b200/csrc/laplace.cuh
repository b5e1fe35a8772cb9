// laplace.cuh -- device kernels of the Laplace Newton loop (filled in below agp.cu's SVGP path).
#pragma once
#include "dense.cuh"
namespace agp {}
