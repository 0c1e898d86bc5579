/**
 * @file operators.hxx
 * @brief Umbrella for the frontier operators (same include path as the reference's, so `#include
 * <gunrock/framework/operators/operators.hxx>` keeps working). Order matters only in that configs.hxx comes first;
 * the advance header pulls in its kernels (kernels.cuh, pull.cuh, directional.cuh, near_far.cuh) itself.
 */
#pragma once

#include <gunrock/framework/operators/configs.hxx>

// traversal
#include <gunrock/framework/operators/advance/advance.hxx>
#include <gunrock/framework/operators/advance/near_far.cuh>
#include <gunrock/framework/operators/neighborreduce/neighborreduce.hxx>
// frontier -> frontier
#include <gunrock/framework/operators/filter/filter.hxx>
#include <gunrock/framework/operators/uniquify/uniquify.hxx>
// 1-D partitioned runs: frontier -> owners
#include <gunrock/framework/operators/exchange/exchange.hxx>
// maps and host-side composition
#include <gunrock/framework/operators/for/for.hxx>
#include <gunrock/framework/operators/batch/batch.hxx>
