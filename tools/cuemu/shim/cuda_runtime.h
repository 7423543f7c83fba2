// Stand-in for <cuda_runtime.h> in the CPU emulation build (tools/cuemu; tests only).
#pragma once
#include "../cuemu.h"
