// forwarding header: the reference's caller includes "path_tracer.h"; everything lives in pt_facade.hpp
#pragma once
#include "pt_facade.hpp"
