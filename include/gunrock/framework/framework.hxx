/**
 * @file framework.hxx
 * @brief Umbrella of the BSP framework: frontier storage, problem/enactor, operators (the include path the
 * reference's algorithm headers use, framework/framework.hxx).
 */
#pragma once

#include <gunrock/memory.hxx>
#include <gunrock/b200/warp.cuh>
#include <gunrock/b200/vector_ops.cuh>
#include <gunrock/framework/frontier/frontier.hxx>
#include <gunrock/framework/operators/operators.hxx>
#include <gunrock/framework/enactor.hxx>
#include <gunrock/framework/problem.hxx>
