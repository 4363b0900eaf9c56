// host_shim/GLFW/glfw3.h — the handful of GLFW names src/IO/*.hpp and src/Camera.cpp mention.  TEST INFRASTRUCTURE ONLY.
#pragma once
struct GLFWwindow;
#define GLFW_MOUSE_BUTTON_LAST 7
#define GLFW_KEY_W 87
#define GLFW_KEY_S 83
#define GLFW_KEY_A 65
#define GLFW_KEY_D 68
#define GLFW_KEY_R 82
#define GLFW_KEY_E 69
#define GLFW_KEY_LEFT_SHIFT 340
#define GLFW_KEY_RIGHT_SHIFT 344
