/** @file cuda.hxx  Umbrella for the CUDA runtime shim (reference: include/gunrock/cuda/cuda.hxx). */
#pragma once
#include <cstdio>
#include <gunrock/cuda/context.hxx>
#include <gunrock/cuda/launch.hxx>
#include <gunrock/cuda/launch_box.hxx>
