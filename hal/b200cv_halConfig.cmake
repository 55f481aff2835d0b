# OpenCV_HAL package description (see /root/reference/samples/hal/c_hal/config.cmake for the shape OpenCV expects):
#   cmake -DOpenCV_HAL_DIR=<this directory> <opencv source>
get_filename_component(_b200cv_root "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(b200cv_hal_FOUND TRUE)
set(b200cv_hal_VERSION "0.1.0")
set(b200cv_hal_LIBRARIES "${_b200cv_root}/opencv_b200/lib/libb200cv.so")
set(b200cv_hal_HEADERS "b200cv_hal_replacement.hpp")
set(b200cv_hal_INCLUDE_DIRS "${CMAKE_CURRENT_LIST_DIR}" "${_b200cv_root}/include")
