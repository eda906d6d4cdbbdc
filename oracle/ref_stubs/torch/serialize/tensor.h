// Stub: lets the reference *.cu kernel files compile without libtorch.
// Their headers only mention at::Tensor in declarations of host shims we do not build.
#pragma once
namespace at { class Tensor; }
