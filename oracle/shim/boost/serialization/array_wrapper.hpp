/* shim: nothing needed (the reference includes this header only to work around a Boost 1.64 bug) */
#pragma once
