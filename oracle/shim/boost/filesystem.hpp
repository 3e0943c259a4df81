#include "tracy_boost_stubs.hpp"
