// Stub: see torch/serialize/tensor.h in this directory.
#pragma once
namespace at { class Tensor; }
