/** @file operators.hxx  Umbrella for the frontier operators. */
#pragma once
#include <gunrock/framework/operators/configs.hxx>
#include <gunrock/framework/operators/for/for.hxx>
#include <gunrock/framework/operators/advance/advance.hxx>
#include <gunrock/framework/operators/filter/filter.hxx>
#include <gunrock/framework/operators/uniquify/uniquify.hxx>
#include <gunrock/framework/operators/neighborreduce/neighborreduce.hxx>
#include <gunrock/framework/operators/batch/batch.hxx>
