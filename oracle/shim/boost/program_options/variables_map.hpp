#include "options_description.hpp"
