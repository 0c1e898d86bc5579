// TEST INFRASTRUCTURE ONLY: see context.hxx in this directory.
#pragma once
#include "context.hxx"
