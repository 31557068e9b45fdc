// TEST INFRASTRUCTURE ONLY (oracle build). The reference's `timestamp` manipulator
// (timestamp.hpp:9-13) streams boost::posix_time::microsec_clock::local_time(); only the
// log text depends on it.
#ifndef EM2_ORACLE_SHIM_POSIX_TIME_HPP
#define EM2_ORACLE_SHIM_POSIX_TIME_HPP
#include <chrono>
#include <ctime>
#include <cstdio>
#include <ostream>
namespace boost { namespace posix_time {
struct ptime { std::chrono::system_clock::time_point t; };
struct microsec_clock { static ptime local_time() { return ptime{std::chrono::system_clock::now()}; } };
inline std::ostream& operator<<(std::ostream& s, const ptime& p)
{
    const std::time_t tt = std::chrono::system_clock::to_time_t(p.t);
    std::tm tmv;
    localtime_r(&tt, &tmv);
    char buf[64];
    std::strftime(buf, sizeof(buf), "%Y-%b-%d %H:%M:%S", &tmv);
    const long us = long(std::chrono::duration_cast<std::chrono::microseconds>(p.t.time_since_epoch()).count() % 1000000);
    char out[96];
    std::snprintf(out, sizeof(out), "%s.%06ld", buf, us);
    return s << out;
}
}}
#endif
