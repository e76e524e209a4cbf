// svgp.cuh -- K7/K8 sparse variational GP drivers.
#pragma once
#include "common.cuh"
