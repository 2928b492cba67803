#pragma once
#include "assert.h"
