// Stand-in for <boost/filesystem.hpp>: the two predicates PinholeStereoCamera's YAML constructor names.
#pragma once
#include <string>
#include <sys/stat.h>
namespace boost { namespace filesystem {
inline bool exists(const std::string &p) { struct stat s; return ::stat(p.c_str(), &s) == 0; }
inline bool is_regular(const std::string &p) { struct stat s; return ::stat(p.c_str(), &s) == 0 && S_ISREG(s.st_mode); }
}} // namespace boost::filesystem
