#include "../ximgproc.hpp"
