// shim: see oracle/shim/msgs_common.hpp
#pragma once
#include "msgs_common.hpp"
