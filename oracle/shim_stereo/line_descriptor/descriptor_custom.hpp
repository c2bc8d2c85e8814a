#include <line_descriptor_custom.hpp>
