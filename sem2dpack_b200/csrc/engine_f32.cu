// FP32 instantiation of the engine (fields and coefficients in single precision; boundary state stays FP64).
#define S2D_INSTANTIATE_F32
#include "engine.hpp"

template class s2d::Engine<float>;
