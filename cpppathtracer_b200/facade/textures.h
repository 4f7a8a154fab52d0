// forwarding header: the reference's caller includes "textures.h"; everything lives in pt_facade.hpp
#pragma once
#include "pt_facade.hpp"
