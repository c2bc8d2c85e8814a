#include "../features2d.hpp"
