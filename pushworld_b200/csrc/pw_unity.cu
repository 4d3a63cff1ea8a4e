// pw_unity.cu -- the library is compiled as a single translation unit so that __constant__
// tables and templates are shared between the kernels without relocatable device code.
#include "pw_kernels.cu"
#include "pw_render.cu"
#include "pw_abi.cu"
