// empty stand-in: src/web.h includes this header and uses nothing from it
