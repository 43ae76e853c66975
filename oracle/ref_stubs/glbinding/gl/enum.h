#include "gl.h"
