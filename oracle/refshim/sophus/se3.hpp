#pragma once
#include "../mini_sophus.h"
