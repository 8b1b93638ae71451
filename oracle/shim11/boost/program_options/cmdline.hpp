#include <boost/program_options.hpp>
