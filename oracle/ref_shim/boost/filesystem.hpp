// boost::filesystem stand-in: the reference only creates the directory of its map file.
#pragma once
#include <string>
#include <sys/stat.h>
namespace boost {
namespace filesystem {
class path {
   public:
    std::string s;
    path() {}
    path(const std::string& p) : s(p) {}
    path parent_path() const {
        const size_t k = s.find_last_of('/');
        return path(k == std::string::npos ? std::string() : s.substr(0, k));
    }
};
inline bool create_directories(const path& p) {
    if (p.s.empty()) return false;
    std::string acc;
    for (size_t i = 0; i <= p.s.size(); i++)
        if (i == p.s.size() || (p.s[i] == '/' && i > 0)) {
            acc = p.s.substr(0, i);
            ::mkdir(acc.c_str(), 0755);
        }
    return true;
}
}  // namespace filesystem
}  // namespace boost
