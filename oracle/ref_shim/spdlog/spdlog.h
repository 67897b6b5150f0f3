// No-op spdlog stand-in for the reference rebuild (arguments are still evaluated, like the real one
// does at any log level) -- test infrastructure only.
#pragma once
namespace spdlog {
template <typename... A> inline void debug(const char*, A&&...) {}
template <typename... A> inline void info(const char*, A&&...) {}
template <typename... A> inline void warn(const char*, A&&...) {}
template <typename... A> inline void error(const char*, A&&...) {}
}  // namespace spdlog
