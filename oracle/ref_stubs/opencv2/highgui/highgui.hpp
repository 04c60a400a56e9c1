// forwards to the single stand-in header (see ../opencv.hpp)
#pragma once
#include "../opencv.hpp"
