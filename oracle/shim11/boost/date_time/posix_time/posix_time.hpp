// Stand-in for the Boost.Date_time names the reference prints progress lines and ##fileDate with (TEST INFRASTRUCTURE ONLY)
#pragma once
#include <ctime>
#include <cstdio>
#include <string>
namespace boost {
namespace gregorian { struct date { std::tm t; }; inline std::string to_iso_string(date const& d) { char b[32]; std::snprintf(b, sizeof(b), "%04d%02d%02d", d.t.tm_year + 1900, d.t.tm_mon + 1, d.t.tm_mday); return b; } }
namespace posix_time {
struct ptime { std::time_t t; gregorian::date date() const { gregorian::date d; localtime_r(&t, &d.t); return d; } };
struct second_clock { static ptime local_time() { return ptime{std::time(nullptr)}; } };
inline std::string to_simple_string(ptime const& p) { std::tm tmv; localtime_r(&p.t, &tmv); char b[64]; std::strftime(b, sizeof(b), "%Y-%b-%d %H:%M:%S", &tmv); return b; }
}}
