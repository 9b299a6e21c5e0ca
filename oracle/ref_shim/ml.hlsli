// TEST INFRASTRUCTURE -- see ml.h in this directory.  The reference's host code includes "ml.hlsli" (Source/InstanceImpl.h:22)
// for the shader-side half of MathLib; none of it is used by the C++ sources, so this shim is intentionally empty.
#pragma once
