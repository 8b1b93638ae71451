// empty stand-in: the reference header includes this but the functions compiled for the oracle do not use it
