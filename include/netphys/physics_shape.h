// include/netphys/physics_shape.h -- the reference splits its API over physics.h and physics_shape.h; the shim keeps both
// include names so `#include "physics_shape.h"` / `#include "physics.h"` lines of reference-style code resolve.
#pragma once
#include "physics.h"
