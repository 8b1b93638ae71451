#include <boost/date_time/posix_time/posix_time.hpp>
