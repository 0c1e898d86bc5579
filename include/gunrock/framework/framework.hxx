/** @file framework.hxx  Umbrella: problem, enactor, frontier, operators (reference framework/framework.hxx). */
#pragma once
#include <gunrock/framework/problem.hxx>
#include <gunrock/framework/enactor.hxx>
#include <gunrock/framework/frontier/frontier.hxx>
#include <gunrock/framework/operators/operators.hxx>
