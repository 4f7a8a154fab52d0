// forwarding header: the reference's caller includes "motional_camera.h"; everything lives in pt_facade.hpp
#pragma once
#include "pt_facade.hpp"
