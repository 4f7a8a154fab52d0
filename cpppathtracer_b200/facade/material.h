// forwarding header: the reference's caller includes "material.h"; everything lives in pt_facade.hpp
#pragma once
#include "pt_facade.hpp"
