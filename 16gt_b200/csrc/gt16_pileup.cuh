// gt16_pileup.cuh -- declarations shared by gt16_pileup.cu and its callers (the types are part of the C ABI, include/gt16.h)
#pragma once
#include <stdint.h>
#include "../../include/gt16.h"
