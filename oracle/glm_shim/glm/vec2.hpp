#pragma once
#include "shim_types.hpp"
