/* empty stand-in: the reference includes SFML here without using it (sdr/IRSDR.cpp:5-6) */
