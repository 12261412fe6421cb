// The reference's generated objective for the ling model of inttest/codegeneration, compiled where it lies.
#include "/root/reference/inttest/codegeneration/ling.cpp"
#include "driver.inc"
