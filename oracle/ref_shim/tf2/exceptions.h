#pragma once
#include <stdexcept>
#include <string>
namespace tf2 {
class TransformException : public std::runtime_error {
   public:
    explicit TransformException(const std::string& what) : std::runtime_error(what) {}
};
class LookupException : public TransformException {
   public:
    explicit LookupException(const std::string& what) : TransformException(what) {}
};
}  // namespace tf2
