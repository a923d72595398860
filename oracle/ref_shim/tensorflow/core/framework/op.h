// See op_kernel.h in this directory — TEST INFRASTRUCTURE ONLY (TensorFlow header stand-in).
#pragma once
#include "tensorflow/core/framework/op_kernel.h"
