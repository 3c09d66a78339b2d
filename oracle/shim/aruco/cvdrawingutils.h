// TEST INFRASTRUCTURE: nothing from aruco's drawing helpers is used by main_vignetteCalib.cpp.
#pragma once
