// stand-in for <boost/shared_ptr.hpp> (test infrastructure): the standard shared_ptr under boost's name
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; using std::dynamic_pointer_cast; using std::make_shared; }
