/*
 * oracle/refshim/img_completion.h -- TEST INFRASTRUCTURE ONLY.
 * The reference includes "img_completion.h" (src/DC_lidar_only/img_completion.cpp:15, main.cpp:1, utils.cpp:3) but the
 * file is not in its repository.  This is the minimal header that lets img_completion.cpp compile unmodified: the
 * includes it evidently relied on and the declaration of the one function it defines (img_completion.cpp:17-20).
 */
#ifndef DCMT_REFSHIM_IMG_COMPLETION_H
#define DCMT_REFSHIM_IMG_COMPLETION_H
#include <iostream>
#include <string>
#include <opencv2/opencv.hpp>

void img_completion(const cv::Mat &sparse_r_img, cv::Mat &dense_r_img, const bool &extr, const std::string &blur_type);
#endif
