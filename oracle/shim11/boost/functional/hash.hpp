#include <boost/filesystem.hpp>
