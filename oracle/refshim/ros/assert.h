// stand-in for <ros/assert.h> / console macros used by the factor sources (test infrastructure)
#pragma once
#include <cassert>
#include <cstdio>
#include <cstdlib>
#define ROS_ASSERT(c) assert(c)
#define ROS_ASSERT_MSG(c, ...) assert(c)
#define ROS_BREAK() abort()
#define ROS_INFO(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
