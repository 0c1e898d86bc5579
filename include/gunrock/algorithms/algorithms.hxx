/**
 * @file algorithms.hxx
 * @brief Umbrella header every algorithm includes (same path as the reference's
 * include/gunrock/algorithms/algorithms.hxx:17-40). Pulls in the operator API and the Thrust headers
 * the reference's algorithm sources rely on transitively, so those sources compile against this tree.
 */
#pragma once

#include <cstdio>
#include <iostream>
#include <limits>
#include <memory>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>
#include <thrust/device_ptr.h>
#include <thrust/fill.h>
#include <thrust/copy.h>
#include <thrust/transform.h>
#include <thrust/transform_reduce.h>
#include <thrust/functional.h>
#include <thrust/sort.h>
#include <thrust/sequence.h>
#include <thrust/reduce.h>
#include <thrust/logical.h>
#include <thrust/extrema.h>
#include <thrust/iterator/counting_iterator.h>

#include <gunrock/memory.hxx>
#include <gunrock/error.hxx>
#include <gunrock/util/math.hxx>
#include <gunrock/util/type_limits.hxx>
#include <gunrock/util/load_store.hxx>
#include <gunrock/cuda/cuda.hxx>
#include <gunrock/graph/graph.hxx>
#include <gunrock/framework/framework.hxx>
#include <gunrock/b200/vector_ops.cuh>
