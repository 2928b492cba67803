#pragma once
#include "../ros_stub.h"
