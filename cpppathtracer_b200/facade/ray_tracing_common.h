// forwarding header: the reference's caller includes "ray_tracing_common.h"; everything lives in pt_facade.hpp
#pragma once
#include "pt_facade.hpp"
