// The reference's generated objective for the multiple-substocks model, compiled where it lies.
#include "/root/reference/baseline/multiple-substocks.cpp"
#include "driver.inc"
