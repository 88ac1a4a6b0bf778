// glm shim (test infrastructure): gtx/fast_exponential — nothing on the OMP path uses it.
#pragma once
#include "../glm.hpp"
