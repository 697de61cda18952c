// point.hpp -- same include name as the reference; the class lives in svo_host.hpp
#pragma once
#include "svo_host.hpp"
