"""bmirobot environments backed by the sm_100a articulated-body kernel (csrc/physics.cu)."""
