// ORACLE / TEST INFRASTRUCTURE ONLY (see THC.h in this directory): atomicAdd(float*, float) is a CUDA builtin.
#pragma once
