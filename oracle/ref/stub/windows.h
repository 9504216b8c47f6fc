/* empty stand-in: the reference's physics_shape.cpp includes <windows.h> only for OpenGL */
#pragma once
