// ros stand-in (oracle/ref_shim/README.md): time, parameters from a table, publishers that keep the last message.
#pragma once
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <typeindex>
#include <vector>

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)

namespace ros {
struct Duration {
    double s = 0.0;
    Duration() {}
    explicit Duration(double sec) : s(sec) {}
    double toSec() const { return s; }
};
struct Time {
    double s = 0.0;
    Time() {}
    explicit Time(double sec) : s(sec) {}
    double toSec() const { return s; }
    static Time& clock() {
        static Time t;
        return t;
    }
    static Time now() { return clock(); }  // the harness advances the clock
    Time operator+(const Duration& d) const { return Time(s + d.s); }
    Duration operator-(const Time& o) const { return Duration(s - o.s); }
    bool operator==(const Time& o) const { return s == o.s; }
};

// last message of every type that went through any publisher (the harness reads them back)
struct Capture {
    std::map<std::type_index, std::shared_ptr<void>> last;
    std::map<std::type_index, long> count;
    static Capture& get() {
        static Capture c;
        return c;
    }
    template <class M>
    void put(const M& m) {
        last[std::type_index(typeid(M))] = std::make_shared<M>(m);
        count[std::type_index(typeid(M))]++;
    }
    template <class M>
    const M* peek() const {
        auto it = last.find(std::type_index(typeid(M)));
        return it == last.end() ? nullptr : static_cast<const M*>(it->second.get());
    }
    template <class M>
    long n() const {
        auto it = count.find(std::type_index(typeid(M)));
        return it == count.end() ? 0 : it->second;
    }
};

struct Publisher {
    template <class M>
    void publish(const M& m) const {
        Capture::get().put(m);
    }
};
struct ServiceServer {};

class NodeHandle {
   public:
    std::map<std::string, std::string> str;
    std::map<std::string, double> num;
    std::map<std::string, std::vector<double>> vec;
    template <class T>
    Publisher advertise(const std::string&, int, bool = false) {
        return Publisher();
    }
    template <class Obj, class Req, class Res>
    ServiceServer advertiseService(const std::string&, bool (Obj::*)(Req&, Res&), Obj*) {
        return ServiceServer();
    }
    template <class T>
    void param(const std::string& name, T& out, const T& def) const {
        auto it = num.find(name);
        out = it == num.end() ? def : static_cast<T>(it->second);
    }
    bool getParam(const std::string& name, std::vector<double>& out) const {
        auto it = vec.find(name);
        if (it == vec.end()) return false;
        out = it->second;
        return true;
    }
};
template <>
inline void NodeHandle::param<std::string>(const std::string& name, std::string& out, const std::string& def) const {
    auto it = str.find(name);
    out = it == str.end() ? def : it->second;
}
template <>
inline void NodeHandle::param<bool>(const std::string& name, bool& out, const bool& def) const {
    auto it = num.find(name);
    out = it == num.end() ? def : it->second != 0.0;
}
}  // namespace ros
