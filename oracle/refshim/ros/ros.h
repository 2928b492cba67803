#pragma once
#include "assert.h"
#include "../ros_stub.h"
